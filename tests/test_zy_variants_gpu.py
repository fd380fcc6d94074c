"""CUDA parity on the parameter-rule / solvent variants of common.variant_systems(): arithmetic combination rule
(GEOM = false kernels), a three-site solvent with LJ on the hydrogens (SPC = false: nonbond_ww / nonbond_qw general
routines), [el_scale] pairs, qq_use_library_charges.  None of the reference's shipped inputs exercises them.
Plus C5 at full size against the oracle's committed fingerprint (until now C5 was covered by properties only).

On the CPU the product's host tables for these systems are checked against the oracle entry by entry
(tests/test_host_tables_cpu.py) and the oracle's gradient against finite differences.  All cases passed on the driver's
B200 at the end of round 1 (GPUTEST_r01.json), so they are ordinary tests now.
"""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", common.variant_systems(), ids=lambda c: c[0])
def test_variant_parity(case):
    from test_parity_gpu import _run_case
    name, q, cuts, lam = case
    _run_case(q, cuts, np.array(lam))


def test_c5_full_size_against_oracle_fingerprint():
    """C5 at BASELINE.json's full size (98 304 atoms, periodic) against the oracle's committed fingerprint
    (tests/golden/c5_box_fingerprint.npz, make_golden.py: the reference's O(ncgp^2) list build takes ~25 s on the CPU, too
    long for a GPU test): list size and an order-independent checksum of the 21.5 M water-water entries, energies, the
    gradient of 4 096 sampled atoms and its global norm, LRF moments of 512 sampled charge groups."""
    import os
    import sys
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    sys.path.insert(0, common.GOLDEN)
    from make_golden import list_checksum
    z = np.load(os.path.join(common.GOLDEN, "c5_box_fingerprint.npz"))
    q, cuts, lam = synth.config("C5")
    assert np.array_equal(z["cuts"], [cuts[k] for k in common.CUT_KEYS])
    g = Qnb(q)
    try:
        c = g.make_pair_lists(q.xtop, **cuts)
        assert np.array_equal(c[:5], z["counts"][:5])
        ij, _ = g.export_list(2, 1, params=False)
        assert np.array_equal(list_checksum(ij), z["sum_ww"])
        del ij
        d, E, EQ = g.pot_energy_nonbonds(q.xtop, lam)
        for k in range(7):
            common.assert_energy(f"E[{k}]", E[k], z["E"][k])
        assert common.rel_rms(d[z["atoms"]], z["d_sample"]) <= common.FORCE_REL_RMS
        assert abs((d ** 2).sum() - float(z["d_norm2"])) <= 2e-5 * float(z["d_norm2"])
        lrf, want = g.export_lrf()[z["groups"]], z["lrf_sample"]
        scale = np.abs(want).max(axis=0) + 1e-300
        tol = np.where(np.arange(43) < 16, 1e-9, 2e-5)
        assert np.all(np.abs(lrf - want) <= tol * scale + 1e-12)
    finally:
        g.close()
