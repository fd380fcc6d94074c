"""Batched windows: where the time of a batched step goes (design experiment): python tools/exp_batch.py C2|C3 [W]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth
from q6_b200.engine import Qnb, QnbBatch
w = sys.argv[1] if len(sys.argv) > 1 else "C2"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 7
q, cuts, lam = synth.config(w)
hs = [Qnb(q) for _ in range(W)]
b = QnbBatch(hs)
rng = np.random.default_rng(1)
xs = [q.xtop + rng.normal(0, 0.01, q.xtop.shape) for _ in range(W)]
b.make_pair_lists(xs, **cuts)
b.pot_energy_nonbonds(xs, [lam] * W)
for _ in range(30):
    b.pot_energy_nonbonds()
t0 = time.perf_counter(); n = 200
for _ in range(n):
    b.pot_energy_nonbonds()
t_step = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for _ in range(5):
    b.make_pair_lists(xs, **cuts)
t_build = (time.perf_counter() - t0) / 5
print(w, "W", W, "batched step us %.1f (per window %.1f)" % (t_step * 1e6, t_step * 1e6 / W), "batched build ms %.3f" % (t_build * 1e3),
      {k: round(v * 1e6, 1) for k, v in b.last_timing().items()})
g = hs[0]
x = q.xtop.copy()
t0 = time.perf_counter()
for _ in range(n):
    g.pot_energy_nonbonds(x, lam)
print("   single window e2e step us %.1f" % ((time.perf_counter() - t0) / n * 1e6), {k: round(v * 1e6, 1) for k, v in g.last_timing().items()})
