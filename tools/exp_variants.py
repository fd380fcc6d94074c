"""Time kernel variants built with -DQNB_EXP_* (design experiments; not part of the product)."""
import sys, os, shutil
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
q, cuts, lam = synth.config(os.environ.get("WL", "C2"))
for v in sys.argv[1:]:
    lib = engine.load_library() if v == "MAIN" else engine.load_library(os.path.abspath(f"tools/exp/libqnb_{v}.so"))
    g = engine.Qnb(q, lib=lib)
    g.make_pair_lists(q.xtop, **cuts, counts=False)
    g.pot_energy_nonbonds(q.xtop, lam)
    kt = g.bench_kernels(lam, 30)
    print(v, {k: round(t * 1e3, 1) for k, t in kt.items()}, "md us/step", round(g.bench_md(lam, 100, 25) * 10, 1))
    g.close()
