"""Is the batched step bound by launches or by work?  Same batch size on a tiny system (design experiment)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth
from q6_b200.engine import Qnb, QnbBatch
W = int(sys.argv[1]) if len(sys.argv) > 1 else 7
R = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
q = synth.solvated_sphere(R, R * 0.5, 10, 1, 5)
cuts = dict(Rq=99.0, Rcq2=99.0 ** 2, RcLRF2=99.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=99.0)
lam = np.array([1.0])
hs = [Qnb(q) for _ in range(W)]
b = QnbBatch(hs)
xs = [q.xtop.copy() for _ in range(W)]
b.make_pair_lists(xs, **cuts)
b.pot_energy_nonbonds(xs, [lam] * W)
for _ in range(30):
    b.pot_energy_nonbonds()
t0 = time.perf_counter(); n = 300
for _ in range(n):
    b.pot_energy_nonbonds()
t = (time.perf_counter() - t0) / n
print("natom", q.natom, "W", W, "batched step us %.1f (per window %.1f)" % (t * 1e6, t * 1e6 / W), {k: round(v * 1e6, 1) for k, v in b.last_timing().items()},
      "launches/step/window", hs[0].launch_count())
