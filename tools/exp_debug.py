"""Debug helper: compare one small case against the oracle, per atom class (design experiment)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from q6_b200 import synth, engine
from oracle.pyoracle import Oracle
import common
name, q, cuts, lam = common.small_systems()[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
lam = np.array(lam)
lib = engine.load_library(os.path.abspath(os.environ['QLIB'])) if os.environ.get('QLIB') else engine.load_library()
g, o = engine.Qnb(q, lib=lib), Oracle(q)
g.make_pair_lists(q.xtop, **cuts); o.make_pair_lists(q.xtop, **cuts)
dg, Eg, EQg = g.pot_energy_nonbonds(q.xtop, lam)
do, Eo, EQo = o.pot_energy_nonbonds(q.xtop, lam)
err = np.abs(dg - do).max(axis=1)
bad = np.where(~(err < 1e-3))[0]
print(name, "natom", q.natom, "nat_solute", q.nat_solute, "bad atoms", len(bad), bad[:20])
print("finite", np.isfinite(dg).all(), "E gpu", Eg, "\nE ora", Eo)
for a in bad[:5]: print(a, dg[a], do[a])
