"""QCP bead loop: batched entry against one qnb_nonbond(md=.false.) call per bead (design experiment)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
q, cuts, lam = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C4s")
g = engine.Qnb(q)
g.make_pair_lists(q.xtop, **cuts, counts=False)
rng = np.random.default_rng(1)
atoms = np.asarray(q.iqseq[:3])
nb = 32
coord = rng.normal(0, 0.05, (nb, 3, 3))
g.qcp_beads(q.xtop, atoms, coord, lam)
t = time.perf_counter()
for _ in range(20): EQ = g.qcp_beads(q.xtop, atoms, coord, lam)
tb = (time.perf_counter() - t) / 20
x = q.xtop.copy()
g.pot_energy_nonbonds(x, lam, md=False)
t = time.perf_counter()
for _ in range(20):
    for b in range(nb):
        xb = x.copy(); xb[atoms - 1] += coord[b]
        g.pot_energy_nonbonds(xb, lam, md=False)
tl = (time.perf_counter() - t) / 20
print(f"{nb} beads natom {q.natom} nqat {q.nqat}: batched {tb*1e3:.3f} ms ({tb/nb*1e6:.1f} us/bead), per-bead calls {tl*1e3:.3f} ms ({tl/nb*1e6:.1f} us/bead)")
