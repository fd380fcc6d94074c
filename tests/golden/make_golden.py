"""Generates tests/golden/*.npz from the reference's shipped test inputs (run in the build
container, where /root/reference exists; the GPU box only sees the committed .npz files).

For each fixture: the parsed system tables (q6_b200.topo/fep/system) and the oracle's results at the
topology coordinates -- list sizes, an order-independent checksum of every list, energies, the
gradient and the LRF moments.  The oracle itself is pinned on the reference's own golden scalars
(test_oracle_golden.py): 56 402 water pairs and E%ww%vdw = -413.17 for prep_SPH.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from q6_b200.fep import load_fep  # noqa: E402
from q6_b200.system import build_system  # noqa: E402
from q6_b200.topo import topo_read  # noqa: E402

REF = "/root/reference/tests"
OUT = os.path.dirname(os.path.abspath(__file__))

FIXTURES = {
    # name: (top, fep, cut-offs as the tests' inputs set them, lambdas)
    # run_test.sh:L171-186: SPH -> q_atom 99, lrf 99; PBC -> q_atom 24, lrf 24; others default 10
    "c1_sph": (f"{REF}/basic_tests/prep_SPH/lig_w.top", f"{REF}/basic_tests/prep_SPH/lig_w.fep",
               dict(Rq=99.0, Rcq2=99.0 ** 2, RcLRF2=99.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=99.0), [1.0]),
    "c1_pbc": (f"{REF}/basic_tests/prep_PBC/lig_w.top", f"{REF}/basic_tests/prep_PBC/lig_w.fep",
               dict(Rq=24.0, Rcq2=24.0 ** 2, RcLRF2=24.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=24.0), [1.0]),
    # exclude_tests/inputs/excl/gen_inps.pl: cut-offs 10/10/10, q_atom 99, lrf on (default 99)
    "c4_evb": (f"{REF}/exclude_tests/inputs/2cjpFH_ionres_oplsa.top", f"{REF}/exclude_tests/inputs/lig.fep",
               dict(Rq=99.0, Rcq2=99.0 ** 2, RcLRF2=99.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=99.0), [0.5, 0.5]),
}


# Row 1 of tests/basic_tests/{SPH,PBC}_*_benchmark.en (lower == upper bound; identical in all six files of a kind):
# QEL QVdW (= "Q-surr." el, vdw: qp + qw) and EL VdW (= solute + solvent + solute-solvent el, vdw; LRF not included),
# eval_test.sh:99-101, 191-196.  Printed with two decimals.
STEP0 = {"c1_sph": (3.12, 139.43, -7.30, -413.17), "c1_pbc": (-31.30, 228.64, 380.33, -1990.55)}


def list_checksum(ij: np.ndarray) -> np.ndarray:
    """Order-independent fingerprint of a pair list: count, sum of i*P+j mod 2^61-1 and xor."""
    if len(ij) == 0:
        return np.array([0, 0, 0], np.int64)
    key = ij[:, 0].astype(np.int64) * 1000003 + ij[:, 1].astype(np.int64)
    mixed = (key * 0x9E3779B97F4A7C15 % (2 ** 61 - 1)) if False else key
    return np.array([len(ij), int(np.sum(mixed % 2147483647)), int(np.bitwise_xor.reduce(key))], np.int64)


def main():
    for name, (top, fep, cuts, lam) in FIXTURES.items():
        t = topo_read(top)
        f = load_fep(fep, t)
        q = build_system(t, f, use_LRF=True)
        q.save(os.path.join(OUT, f"{name}_system.npz"))
        o = Oracle(q)
        counts = o.make_pair_lists(q.xtop, **cuts)
        d, E, EQ = o.pot_energy_nonbonds(q.xtop, lam)
        res = dict(counts=counts, E=E, EQ=EQ, d=d, lrf=o.export_lrf(), lam=np.array(lam),
                   cuts=np.array([cuts[k] for k in ("Rq", "Rcq2", "RcLRF2", "Rcpp2", "Rcpw2", "Rcww2", "RcLRF")]))
        for which, nm in enumerate(("pp", "pw", "ww", "qp", "qw", "qq", "qqp")):
            ij, p = o.export_list(which, 1)
            res[f"sum_{nm}"] = list_checksum(ij)
        if name in STEP0:
            # the coordinates of the reference's step 0: topology coordinates after the initial solvent SHAKE
            # (qdyn.f90:133 -> initial_constraint, bondene.f90:1025-1048), and the oracle's results there
            xs, nits = pyoracle.initial_constraint_x(t)
            res["x_step0"] = xs
            res["shake_iterations"] = np.array(nits, np.int32)
            res["counts_step0"] = o.make_pair_lists(xs, **cuts)
            d0, E0, EQ0 = o.pot_energy_nonbonds(xs, lam)
            res.update(d_step0=d0, E_step0=E0, EQ_step0=EQ0)
            print(name, "step 0 (post-SHAKE): QEL %.4f QVdW %.4f EL %.4f VdW %.4f" % (
                EQ0[0, 2] + EQ0[0, 4], EQ0[0, 3] + EQ0[0, 5], E0[0] + E0[2] + E0[4], E0[1] + E0[3] + E0[5]),
                "reference:", STEP0[name])
        np.savez_compressed(os.path.join(OUT, f"{name}_oracle.npz"), **res)
        print(name, "natom", q.natom, "counts", counts[:5], "E", np.round(E, 3), "EQ", np.round(EQ, 3))


def full_size_fingerprints():
    """Oracle results at BASELINE.json's full sizes for the configurations the oracle cannot be run on inside a GPU test
    in seconds (C5: 98 304 atoms, the O(ncgp^2) reference list build takes ~25 s): list sizes, an order-independent
    checksum of the water-water list, energies, the gradient of a fixed random sample of atoms plus its global norm, and
    the LRF moments of a sample of charge groups.  No reference files needed (synthetic system)."""
    from q6_b200 import synth
    q, cuts, lam = synth.config("C5")
    o = Oracle(q)
    counts = o.make_pair_lists(q.xtop, **cuts)
    d, E, EQ = o.pot_energy_nonbonds(q.xtop, lam)
    rng = np.random.default_rng(20261018)
    atoms = np.sort(rng.choice(q.natom, 4096, replace=False))
    groups = np.sort(rng.choice(q.ncgp, 512, replace=False))
    ij, _ = o.export_list(2, 1, params=False)
    np.savez_compressed(os.path.join(OUT, "c5_box_fingerprint.npz"), counts=counts, E=E, EQ=EQ, atoms=atoms, d_sample=d[atoms],
                        d_norm2=float((d ** 2).sum()), d_abs_sum=float(np.abs(d).sum()), groups=groups,
                        lrf_sample=o.export_lrf()[groups], sum_ww=list_checksum(ij),
                        cuts=np.array([cuts[k] for k in ("Rq", "Rcq2", "RcLRF2", "Rcpp2", "Rcpw2", "Rcww2", "RcLRF")]))
    print("C5 fingerprint: counts", counts[:5], "E", np.round(E, 3))


if __name__ == "__main__":
    main()
    full_size_fingerprints()
