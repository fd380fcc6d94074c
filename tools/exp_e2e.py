"""Where does the end-to-end time go? (design experiment)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
q, cuts, lam = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C2")
g = engine.Qnb(q, device=int(os.environ.get('DEV', '0')))
x = q.xtop.copy(); d = np.zeros((q.natom, 3))
for k in range(30):
    if k % 25 == 0: g.make_pair_lists(x, **cuts, counts=False)
    g.pot_energy_nonbonds(x, lam, d=d)
t0 = time.perf_counter()
for k in range(20): g.make_pair_lists(x, **cuts, counts=False)
tb = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
for k in range(400): g.pot_energy_nonbonds(x, lam, d=d)
ts = (time.perf_counter() - t0) / 400
# first step after a build (graph capture + instantiate)
tf = 0
for k in range(10):
    g.make_pair_lists(x, **cuts, counts=False)
    t0 = time.perf_counter(); g.pot_energy_nonbonds(x, lam, d=d); tf += time.perf_counter() - t0
print("graph=%s" % (os.environ.get("QNB_NO_GRAPH", "0") != "1"), "build e2e us: %.1f  step e2e us: %.1f  first step after build us: %.1f" % (tb * 1e6, ts * 1e6, tf / 10 * 1e6), g.last_timing())
