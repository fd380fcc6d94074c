"""Shared helpers of the parity tests."""
import os

import numpy as np

from q6_b200 import synth
from q6_b200.system import QSystem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CUT_KEYS = ("Rq", "Rcq2", "RcLRF2", "Rcpp2", "Rcpw2", "Rcww2", "RcLRF")
LIST_NAMES = ("pp", "pw", "ww", "qp", "qw", "qq", "qqp")

# parity bars of BASELINE.json north_star
FORCE_REL_RMS = 1e-5
ENERGY_REL = 1e-6


def golden_system(name):
    q = QSystem.load(os.path.join(GOLDEN, f"{name}_system.npz"))
    z = np.load(os.path.join(GOLDEN, f"{name}_oracle.npz"))
    cuts = dict(zip(CUT_KEYS, z["cuts"].tolist()))
    return q, cuts, z["lam"], z


def sph_cuts(rc=10.0, rq=99.0, rlrf=99.0):
    return dict(Rq=rq, Rcq2=rq * rq, RcLRF2=rlrf * rlrf, Rcpp2=rc * rc, Rcpw2=rc * rc, Rcww2=rc * rc, RcLRF=rlrf)


def sorted_pairs(ij):
    if len(ij) == 0:
        return np.zeros(0, np.int64)
    key = ij[:, 0].astype(np.int64) * (1 << 32) + ij[:, 1].astype(np.int64)
    return np.sort(key)


def sorted_pairs_with_params(ij, p):
    if len(ij) == 0:
        return np.zeros(0, np.int64), np.zeros((0, 4))
    key = ij[:, 0].astype(np.int64) * (1 << 32) + ij[:, 1].astype(np.int64)
    o = np.argsort(key, kind="stable")
    return key[o], p[o]


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def assert_energy(name, got, want, tol=ENERGY_REL):
    got, want = np.asarray(got), np.asarray(want)
    err = np.abs(got - want)
    bad = err > tol * np.abs(want) + 1e-9   # 1e-9 kcal/mol floor for terms that are exactly zero
    assert not bad.any(), f"{name}: got {got}, want {want}, rel {err / np.maximum(np.abs(want), 1e-300)}"


def small_systems():
    """(name, system, cut-offs, lambdas) of the synthetic parity cases that the oracle finishes in seconds."""
    out = []
    out.append(("sph_fep2", synth.solvated_sphere(16.0, 9.0, 20, 2, 11, fep="annihilate"), sph_cuts(8.0), [0.6, 0.4]))
    out.append(("sph_evb2", synth.solvated_sphere(16.0, 9.0, 40, 2, 12, fep="evb"), sph_cuts(9.0), [0.35, 0.65]))
    out.append(("sph_1state_nolrf", _nolrf(synth.solvated_sphere(15.0, 8.0, 10, 1, 13)), sph_cuts(8.0), [1.0]))
    out.append(("sph_excl_shell", synth.solvated_sphere(16.0, 9.0, 12, 1, 14, excl_shell=2.5), sph_cuts(8.0, rq=12.0, rlrf=14.0), [1.0]))
    out.append(("sph_small_rcq", synth.solvated_sphere(16.0, 9.0, 12, 2, 15, fep="evb"), sph_cuts(7.0, rq=9.0, rlrf=20.0), [0.5, 0.5]))
    out.append(("sph_noq", synth.solvated_sphere(14.0, 8.0, 0, 1, 16), sph_cuts(8.0), [1.0]))
    out.append(("sph_nowater", _nowater(), sph_cuts(8.0), [1.0]))
    box = 8 * synth.A_LATTICE
    out.append(("box_solute_q", synth.solvated_sphere(0.0, 7.0, 16, 2, 17, fep="evb", pbc_box=box),
                dict(Rq=9.0, Rcq2=81.0, RcLRF2=11.5 ** 2, Rcpp2=64.0, Rcpw2=64.0, Rcww2=64.0, RcLRF=11.5), [0.5, 0.5]))
    out.append(("box_solute_q_nocut", synth.solvated_sphere(0.0, 7.0, 16, 1, 18, pbc_box=box),
                dict(Rq=-1.0, Rcq2=1.0, RcLRF2=1.0, Rcpp2=64.0, Rcpw2=64.0, Rcww2=64.0, RcLRF=-1.0), [1.0]))
    out.append(("box_water", synth.water_box(9, 19), dict(Rq=-1.0, Rcq2=1.0, RcLRF2=12.0 ** 2, Rcpp2=81.0, Rcpw2=81.0, Rcww2=81.0, RcLRF=12.0), [1.0]))
    out.append(("box_water_nolrf", _nolrf(synth.water_box(7, 20)), dict(Rq=-1.0, Rcq2=1.0, RcLRF2=1.0, Rcpp2=49.0, Rcpw2=49.0, Rcww2=49.0, RcLRF=-1.0), [1.0]))
    # boxes large against the LRF cut-off: the LRF scan takes the periodic image per cell row (and skips far rows)
    out.append(("box_water_rowimage", synth.water_box(10, 27), dict(Rq=-1.0, Rcq2=1.0, RcLRF2=8.0 ** 2, Rcpp2=36.0, Rcpw2=36.0, Rcww2=36.0, RcLRF=8.0), [1.0]))
    out.append(("box_solute_rowimage", synth.solvated_sphere(0.0, 6.0, 10, 2, 28, fep="evb", pbc_box=10 * synth.A_LATTICE),
                dict(Rq=8.0, Rcq2=64.0, RcLRF2=8.5 ** 2, Rcpp2=36.0, Rcpw2=30.25, Rcww2=36.0, RcLRF=8.5), [0.4, 0.6]))
    # any-atom charge-group cut-offs (iuse_switch_atom = 0: nb??lis2*)
    out.append(("sph_anyatom", _anyatom(synth.solvated_sphere(16.0, 9.0, 12, 2, 22, fep="evb")), sph_cuts(8.0, rq=10.0, rlrf=14.0), [0.5, 0.5]))
    out.append(("sph_anyatom_lrf99", _anyatom(synth.solvated_sphere(15.0, 9.0, 8, 1, 23)), sph_cuts(7.5), [1.0]))
    out.append(("sph_anyatom_nolrf", _nolrf(_anyatom(synth.solvated_sphere(15.0, 9.0, 8, 1, 24))), sph_cuts(8.0), [1.0]))
    out.append(("box_anyatom", _anyatom(synth.solvated_sphere(0.0, 7.0, 16, 2, 25, fep="evb", pbc_box=box)),
                dict(Rq=9.0, Rcq2=81.0, RcLRF2=11.0 ** 2, Rcpp2=49.0, Rcpw2=49.0, Rcww2=49.0, RcLRF=11.0), [0.5, 0.5]))
    out.append(("box_anyatom_nolrfcut", _anyatom(synth.solvated_sphere(0.0, 7.0, 16, 1, 26, pbc_box=box)),
                dict(Rq=-1.0, Rcq2=1.0, RcLRF2=-1.0, Rcpp2=49.0, Rcpw2=49.0, Rcww2=49.0, RcLRF=-1.0), [1.0]))
    return out


def variant_systems():
    """Parameter-rule and solvent variants that none of the reference's shipped inputs exercises (they are all Q-OPLSAA:
    geometric rule, SPC-type TIP3P): arithmetic combination rule, a three-site solvent with LJ on the hydrogens (either
    one sends pot_energy_nonbonds to the general ww/qw routines, potene.f90:347), [el_scale] pairs and
    qq_use_library_charges."""
    box = 8 * synth.A_LATTICE
    out = []
    out.append(("sph_arith_evb", synth.solvated_sphere(16.0, 9.0, 20, 2, 31, fep="evb", ivdw_rule=2), sph_cuts(8.0), [0.4, 0.6]))
    out.append(("sph_general_solvent", synth.solvated_sphere(15.0, 8.0, 12, 1, 32, solvent_type=1), sph_cuts(8.0), [1.0]))
    out.append(("box_arith_general", synth.solvated_sphere(0.0, 7.0, 16, 2, 33, fep="evb", pbc_box=box, ivdw_rule=2, solvent_type=1),
                dict(Rq=9.0, Rcq2=81.0, RcLRF2=11.5 ** 2, Rcpp2=64.0, Rcpw2=64.0, Rcww2=64.0, RcLRF=11.5), [0.5, 0.5]))
    out.append(("sph_elscale_libcrg", synth.solvated_sphere(16.0, 9.0, 20, 2, 34, fep="evb", el_scale=True,
                                                            qq_use_library_charges=True), sph_cuts(8.0), [0.3, 0.7]))
    out.append(("sph_arith_fep", synth.solvated_sphere(15.0, 8.0, 14, 2, 35, fep="annihilate", ivdw_rule=2, el_scale=True),
                sph_cuts(8.0), [0.6, 0.4]))
    return out


def _anyatom(q):
    q.iuse_switch_atom = 0
    return q


def _nolrf(q):
    q.use_LRF = 0
    return q


def _nowater():
    q = synth.solvated_sphere(9.0, 8.5, 14, 1, 21)
    # drop the waters: a solute-only system (natom == nat_solute)
    ns = q.nat_solute
    q.natom, q.nwat = ns, 0
    q.ncgp = q.ncgp_solute
    q.cgp = q.cgp[:q.ncgp]
    q.cgpatom = q.cgpatom[:ns]
    for k in ("iac", "crg", "excl", "iqatom"):
        setattr(q, k, getattr(q, k)[:ns])
    q.xtop = q.xtop[:ns]
    return q.full_shard()
