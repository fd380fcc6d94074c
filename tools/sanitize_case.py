"""Named parity cases (tests/common.py) through list build + two steps + a rebuild, without the oracle; run under
compute-sanitizer by tools/gpu_sanitize.sh.  The parity tests check the numbers, this checks the memory accesses."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from q6_b200.engine import Qnb  # noqa: E402

want = sys.argv[1:] or ["sph_evb2", "sph_small_rcq", "box_solute_q", "box_water", "box_water_rowimage", "box_solute_rowimage", "box_anyatom"]
for name, q, cuts, lam in common.small_systems():
    if name not in want and want != ["all"]:
        continue
    g = Qnb(q, device=0)
    lam = np.array(lam)
    for it in range(2):
        c = g.make_pair_lists(q.xtop + 0.01 * it, **cuts)
        for s in range(2):
            d, E, EQ = g.pot_energy_nonbonds(q.xtop + 0.01 * it + 0.001 * s, lam)
    print(name, c[:5].tolist(), float(np.asarray(E)[:7].sum()), "launches", g.launch_count(), flush=True)
    g.close()
