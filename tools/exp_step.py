"""Step-only timing under graph / stream variants (design experiment)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
q, cuts, lam = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C2")
g = engine.Qnb(q)
g.make_pair_lists(q.xtop, **cuts, counts=False)
g.pot_energy_nonbonds(q.xtop, lam)
g.bench_nonbond(lam, 50)
print("graph=%s one_stream=%s" % (os.environ.get("QNB_NO_GRAPH", "0") != "1", os.environ.get("QNB_ONE_STREAM", "0")),
      "step-only us:", round(g.bench_nonbond(lam, 400) / 400 * 1e3, 2), " build ms:", round(g.bench_build_lists(5) / 5, 3),
      " step without pp/pw/ww energies us:", round(g.bench_nonbond(lam, 400, energies=False) / 400 * 1e3, 2))
if not q.use_PBC and q.nwat > 0:
    rw = float(np.linalg.norm(q.xtop[q.nat_solute:] - np.asarray(q.xpcent), axis=1).max())
    g.set_solvent_restraints(engine.wat_shells(q.xpcent, rw, crgQtot=-1.0), np.zeros(8))
    x = q.xtop.copy(); d = np.zeros((q.natom, 3))
    import time
    print('   device step us with restraints:', round(g.bench_nonbond(lam, 400, restraints=True) / 400 * 1e3, 2), {k: round(v * 1e3, 1) for k, v in g.bench_kernels(lam, 20, restraints=True).items()})
    for r in (False, True):
        for _ in range(50): g.pot_energy_nonbonds(x, lam, d=d, restraints=r)
        t = time.perf_counter()
        for _ in range(500): g.pot_energy_nonbonds(x, lam, d=d, restraints=r)
        print("   e2e step us, solvent restraints in the device step =", r, round((time.perf_counter() - t) / 500 * 1e6, 1))
