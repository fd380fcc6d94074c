"""Device SHAKE of solvent molecules (qnb_set_constraints / qnb_shake) against the CPU restatement of
shake(xx, x), bondene.f90:1069-1150 (oracle.pyoracle.shake; itself pinned by the reference's step-0 goldens, which only
come out right with the reference's flag-once order of operations).

All cases passed on the driver's B200 at the end of round 1 (GPUTEST_r01.json): ordinary tests now.
"""
import os

import numpy as np
import pytest

import common
from common import golden_system

pytestmark = pytest.mark.gpu


def _water_constraints(q):
    from q6_b200 import synth
    return synth.water_constraints(q)


@pytest.mark.parametrize("case", [c for c in common.small_systems() if c[0] in ("sph_fep2", "box_water", "sph_noq")],
                         ids=lambda c: c[0])
def test_shake_matches_cpu_restatement_bit_for_bit(case):
    from oracle import pyoracle
    from q6_b200.engine import Qnb
    name, q, cuts, lam = case
    cons, starts, winv = _water_constraints(q)
    g = Qnb(q)
    try:
        g.set_constraints(cons, starts, winv)
        # (1) initial_constraint: xx = x = jittered rigid waters
        want, nits = pyoracle.shake(cons, starts, winv, q.xtop, q.xtop)
        got, n = g.shake(q.xtop, xx=q.xtop)
        assert n == sum(nits)
        assert np.array_equal(got, want)
        assert np.array_equal(got[:q.nat_solute], q.xtop[:q.nat_solute])
        # (2) the leap-frog case: xx = the coordinates resident from the step's evaluation, x = displaced coordinates
        g.make_pair_lists(want, **cuts)
        g.pot_energy_nonbonds(want, np.array(lam))
        rng = np.random.default_rng(8)
        moved = want + rng.normal(0, 0.01, want.shape)
        want2, nits2 = pyoracle.shake(cons, starts, winv, want, moved)
        got2, n2 = g.shake(moved)
        assert n2 == sum(nits2) and np.array_equal(got2, want2)
    finally:
        g.close()


def test_shake_reference_step0_coordinates():
    """The reference's shipped water sphere: device SHAKE of the topology coordinates == the fixture's post-SHAKE
    coordinates (the ones that reproduce the step-0 goldens)."""
    from q6_b200.engine import Qnb
    q, cuts, lam, z = golden_system("c1_sph")
    mass = np.asarray(q.iaclib).reshape(-1, 7)[np.asarray(q.iac) - 1, 0]
    # lig_w.top: bond codes 13 (O-H, 0.957) and 14 (H-H, 1.5183), tests/basic_tests/prep_SPH
    cons = []
    for k in range(q.nwat):
        o = q.nat_solute + 3 * k + 1
        cons += [(o, o + 1, 0.957 ** 2), (o, o + 2, 0.957 ** 2), (o + 1, o + 2, 1.5183 ** 2)]
    starts = [1] + [q.nat_solute + 3 * k + 1 for k in range(q.nwat)]
    g = Qnb(q)
    try:
        g.set_constraints(cons, starts, 1.0 / mass)
        got, n = g.shake(q.xtop, xx=q.xtop)
        assert np.array_equal(got, z["x_step0"])
        assert n == int(z["shake_iterations"].sum())
    finally:
        g.close()


def test_shake_refuses_solute_sized_molecules():
    from q6_b200.engine import Qnb, QnbError
    q, cuts, lam, z = golden_system("c1_sph")
    g = Qnb(q)
    try:
        cons = [(1, a, 1.0) for a in range(2, 40)]          # 38 constraints in one molecule
        with pytest.raises(QnbError, match="solvent-sized"):
            g.set_constraints(cons, [1], np.ones(q.natom))
    finally:
        g.close()
