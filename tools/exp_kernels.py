"""Per-kernel times and step time of one workload (design experiment): python tools/exp_kernels.py C2|C3|C5 [flush]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
w = sys.argv[1] if len(sys.argv) > 1 else "C2"
q, cuts, lam = synth.config(w)
g = engine.Qnb(q)
c = g.make_pair_lists(q.xtop, **cuts)
g.pot_energy_nonbonds(q.xtop, lam)
g.bench_nonbond(lam, 50)
step = g.bench_nonbond(lam, 400) / 400 * 1e3
noe = g.bench_nonbond(lam, 400, energies=False) / 400 * 1e3
build = g.bench_build_lists(5) / 5
kt = g.bench_kernels(lam, 20, flush_l2=len(sys.argv) > 2)
print(w, "counts", [int(v) for v in c[:5]], "step us %.2f  (no pp/pw/ww energies %.2f)  build ms %.3f" % (step, noe, build),
      {k: round(v, 3) for k, v in g.last_build_timing().items()})
print("   kernels us:", {k: round(v * 1e3, 2) for k, v in kt.items()})
nww = c[2] / 9.0
fl_w = nww * 209 + 0.5 * c[1] * 33
fl_s = c[0] * 33 + 0.5 * c[1] * 33
for k, fl in (("k_water_rows", fl_w), ("k_solute_rows", fl_s)):
    if k in kt and kt[k] > 0:
        print("   %s: %.1f MFLOP algorithmic, %.2f TFLOP/s" % (k, fl / 1e6, fl / (kt[k] * 1e-3) / 1e12))
rows = sum(kt.get(k, 0) for k in ("k_water_rows", "k_solute_rows", "k_pair_energy"))
print("   rows group (forces + energies, serial sum): %.2f TFLOP/s" % ((fl_w + fl_s) / (rows * 1e-3) / 1e12))
