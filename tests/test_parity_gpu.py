"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- the first gate.

Bars (BASELINE.json north_star): pair lists equal as sorted sets (bit-exact indices, parameters to
FP64 round-off), gradient <= 1e-5 relative RMS, every energy term <= 1e-6 relative.
"""
import numpy as np
import pytest

import common
from common import (ENERGY_REL, FORCE_REL_RMS, LIST_NAMES, assert_energy, golden_system, rel_rms, small_systems,
                    sorted_pairs, sorted_pairs_with_params)

pytestmark = pytest.mark.gpu


def _check_lists(g, o, nstates):
    for which, nm in enumerate(LIST_NAMES):
        for st in range(1, nstates + 1 if which >= 3 else 2):
            gi, gp = g.export_list(which, st)
            oi, op = o.export_list(which, st)
            assert len(gi) == len(oi), f"nb{nm} state {st}: {len(gi)} entries, oracle {len(oi)}"
            gk, gpp = sorted_pairs_with_params(gi, gp)
            ok, opp = sorted_pairs_with_params(oi, op)
            assert np.array_equal(gk, ok), f"nb{nm} state {st}: pair sets differ"
            if len(gk):
                assert np.allclose(gpp, opp, rtol=1e-13, atol=1e-300), f"nb{nm} state {st}: parameters differ"


def _check_step(g, o, q, x, lam, md=True, qq=True):
    dg, Eg, EQg = g.pot_energy_nonbonds(x, lam, md=md, qq=qq)
    do, Eo, EQo = o.pot_energy_nonbonds(x, lam, md=md, qq=qq)
    r = rel_rms(dg, do)
    assert r <= FORCE_REL_RMS, f"gradient relative RMS {r:.3e}"
    for k, nm in enumerate(("pp.el", "pp.vdw", "pw.el", "pw.vdw", "ww.el", "ww.vdw", "LRF")):
        assert_energy(nm, Eg[k], Eo[k])
    for s in range(q.nstates):
        for k, nm in enumerate(("qq.el", "qq.vdw", "qp.el", "qp.vdw", "qw.el", "qw.vdw")):
            assert_energy(f"state {s + 1} {nm}", EQg[s, k], EQo[s, k])
    return r


def _run_case(q, cuts, lam, check_lists=True):
    from oracle.pyoracle import Oracle
    from q6_b200.engine import Qnb
    g, o = Qnb(q), Oracle(q)
    try:
        x = q.xtop
        cg = g.make_pair_lists(x, **cuts)
        co = o.make_pair_lists(x, **cuts)
        assert np.array_equal(cg[:5], co[:5]), f"list sizes {cg[:5]} vs oracle {co[:5]}"
        if check_lists:
            _check_lists(g, o, q.nstates)
        if q.use_LRF:
            lg, lo = g.export_lrf(), o.export_lrf()
            scale = np.abs(lo).max(axis=0) + 1e-300
            # centres, phi0, phi1, phi2 are FP64 on the GPU; phi3 (a second-order correction of the field) is FP32
            tol = np.where(np.arange(43) < 16, 1e-9, 2e-5)
            assert np.all(np.abs(lg - lo) <= tol * scale + 1e-12), "LRF moments differ"
        r = _check_step(g, o, q, x, lam)
        # a second evaluation at moved coordinates with the SAME lists (what happens between list updates)
        rng = np.random.default_rng(5)
        x2 = x + rng.normal(0, 0.02, x.shape)
        _check_step(g, o, q, x2, lam)
        # Q-only evaluation (pot_energy(...,.false.) as called by QCP)
        _check_step(g, o, q, x2, lam, md=False)
        # QNB_FLAG_NO_ENERGY: same gradient, LRF and Q terms; the pp/pw/ww energies are simply not produced
        d1, E1, EQ1 = g.pot_energy_nonbonds(x2, lam)
        d1 = d1.copy()
        d0, E0, EQ0 = g.pot_energy_nonbonds(x2, lam, energies=False)
        assert rel_rms(d0, d1) < 1e-9
        assert np.all(E0[:6] == 0.0), E0
        assert np.isclose(E0[6], E1[6], rtol=1e-12, atol=1e-12) and np.allclose(EQ0, EQ1, rtol=1e-12, atol=1e-12)
        return r
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("case", small_systems(), ids=lambda c: c[0])
def test_synthetic_small(case):
    name, q, cuts, lam = case
    _run_case(q, cuts, np.array(lam))


@pytest.mark.parametrize("name", ["c1_sph", "c1_pbc", "c4_evb"])
def test_reference_fixtures(name):
    """The reference's own shipped systems (tests/basic_tests, tests/exclude_tests) at topology coordinates,
    against the live oracle and against the committed oracle results."""
    from q6_b200.engine import Qnb
    q, cuts, lam, z = golden_system(name)
    _run_case(q, cuts, lam)
    g = Qnb(q)
    try:
        c = g.make_pair_lists(q.xtop, **cuts)
        assert np.array_equal(c[:5], z["counts"][:5])
        d, E, EQ = g.pot_energy_nonbonds(q.xtop, lam)
        assert rel_rms(d, z["d"]) <= FORCE_REL_RMS
        for k in range(7):
            assert_energy(f"E[{k}]", E[k], z["E"][k])
        assert_energy("EQ", EQ, z["EQ"])
    finally:
        g.close()


def test_golden_scalars_c1():
    """tests/basic_tests step-0 goldens reachable from topology coordinates: 56 402 water pairs,
    E%ww%vdw = -413.17 (SPH_leap-frog_berendsen_benchmark.en row 1)."""
    from q6_b200.engine import Qnb
    q, cuts, lam, z = golden_system("c1_sph")
    g = Qnb(q)
    try:
        c = g.make_pair_lists(q.xtop, **cuts)
        assert c[2] == 56402 * 9
        d, E, EQ = g.pot_energy_nonbonds(q.xtop, lam)
        assert round(E[5], 2) == -413.17
    finally:
        g.close()


@pytest.mark.parametrize("name", ["c1_sph", "c1_pbc"])
def test_reference_step0_goldens(name):
    """The CUDA path against the reference's OWN numbers: row 1 of tests/basic_tests/{SPH,PBC}_*_benchmark.en and
    eval_test.sh:99-101 (QEL, QVdW, EL, VdW at step 0, printed with two decimals), evaluated at the post-SHAKE
    coordinates of the fixture (qdyn.f90:133), plus the oracle's full result there."""
    from q6_b200.engine import Qnb
    from test_oracle_golden import STEP0_GOLDEN, step0_terms
    q, cuts, lam, z = golden_system(name)
    g = Qnb(q)
    try:
        c = g.make_pair_lists(z["x_step0"], **cuts)
        assert np.array_equal(c[:5], z["counts_step0"][:5])
        d, E, EQ = g.pot_energy_nonbonds(z["x_step0"], lam)
        got = step0_terms(E, EQ)
        for v, want in zip(got, STEP0_GOLDEN[name]):
            assert abs(v - want) <= 0.005 + 1e-9, (got, STEP0_GOLDEN[name])
        assert rel_rms(d, z["d_step0"]) <= FORCE_REL_RMS
        for k in range(7):
            assert_energy(f"E[{k}]", E[k], z["E_step0"][k])
        assert_energy("EQ", EQ, z["EQ_step0"])
    finally:
        g.close()


@pytest.mark.parametrize("name", ["C2", "C3", "C4s"])
def test_baseline_configs(name):
    """The BASELINE.json configurations at full size (synthetic systems of the named shapes)."""
    from q6_b200 import synth
    q, cuts, lam = synth.config(name)
    _run_case(q, cuts, lam, check_lists=True)


@pytest.mark.parametrize("shard_rows", [False, True], ids=["pair_partition", "row_partition"])
@pytest.mark.parametrize("case", ["sphere_q", "water_box", "water_box_shell"])
def test_shards_on_one_gpu_sum_to_the_whole(case, shard_rows, monkeypatch):
    """The sharded list build (the GENERAL instantiations of the LRF kernels, the shard tests of the row scan) without a
    second GPU: every rank's handle is built on cuda:0 with no communicator, so its pair counts and LRF moments stay
    partial, and their sum over the ranks must be the reference's lists and lrf_update (what lrf_gather,
    nonbondene.f90:616, produces).  The all-reduce itself needs >= 2 GPUs: test_multi_gpu.py."""
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    from q6_b200.system import shard_system
    if shard_rows:
        monkeypatch.delenv("QNB_SHARD_PAIRS", raising=False)
    else:
        monkeypatch.setenv("QNB_SHARD_PAIRS", "1")     # read by qnb_init: the reference's pair partition
    if case == "sphere_q":
        q = synth.solvated_sphere(16.0, 9.0, 20, 2, 61, fep="evb")
        cuts = common.sph_cuts(8.0)
    elif case == "water_box":
        q = synth.water_box(10, 62)
        cuts = dict(Rq=-1.0, Rcq2=1.0, RcLRF2=13.0 ** 2, Rcpp2=81.0, Rcpw2=81.0, Rcww2=81.0, RcLRF=13.0)
    else:   # a box large enough for the row-image scan (2(m+1) <= n cells): the kernel the sharded C5 bench runs
        q = synth.water_box(14, 63)
        cuts = dict(Rq=-1.0, Rcq2=1.0, RcLRF2=11.0 ** 2, Rcpp2=49.0, Rcpw2=49.0, Rcww2=49.0, RcLRF=11.0)
    o = Oracle(q)
    want_counts = o.make_pair_lists(q.xtop, **cuts)
    want = o.export_lrf()
    for world in (2, 3):
        counts = np.zeros(5, dtype=np.int64)
        mom = np.zeros_like(want[:, 3:])
        for rank in range(world):
            g = Qnb(shard_system(q, rank, world), device=0)
            try:
                counts += np.asarray(g.make_pair_lists(q.xtop, **cuts))[:5]
                lrf = g.export_lrf()
                assert np.allclose(lrf[:, :3], want[:, :3], rtol=0, atol=1e-9)   # centres: complete on every rank
                mom += lrf[:, 3:]
            finally:
                g.close()
        assert np.array_equal(counts, np.asarray(want_counts)[:5])
        scale = np.abs(want[:, 3:]).max(axis=0) + 1e-300
        tol = np.where(np.arange(3, 43) < 16, 1e-9, 2e-5)
        assert np.all(np.abs(mom - want[:, 3:]) <= tol * scale + 1e-12), (case, world)


def test_list_rebuild_after_motion():
    """Lists rebuilt at moved coordinates equal the oracle's; once-only Q lists stay (nbqplist L3678)."""
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    q = synth.solvated_sphere(16.0, 9.0, 20, 2, 31, fep="evb")
    cuts = common.sph_cuts(8.0)
    g, o = Qnb(q), Oracle(q)
    rng = np.random.default_rng(1)
    x = q.xtop.copy()
    for it in range(3):
        cg, co = g.make_pair_lists(x, **cuts), o.make_pair_lists(x, **cuts)
        assert np.array_equal(cg[:5], co[:5])
        for which in range(5):
            assert np.array_equal(sorted_pairs(g.export_list(which)[0]), sorted_pairs(o.export_list(which)[0]))
        _check_step(g, o, q, x, np.array([0.5, 0.5]))
        x = x + rng.normal(0, 0.15, x.shape)
    g.close()


def test_properties_water_box_full_size():
    """C5 (98 304 atoms, PBC) through size-independent properties: the gradient sums to zero (Newton's third
    law over every listed pair), the pair count is symmetric, and a rigid translation by a lattice vector of
    the box leaves energies unchanged."""
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    q, cuts, lam = synth.config("C5")
    g = Qnb(q)
    try:
        c = g.make_pair_lists(q.xtop, **cuts)
        assert c[2] % 9 == 0 and c[2] > 0
        d, E, EQ = g.pot_energy_nonbonds(q.xtop, lam)
        # ww gradient alone cancels pairwise; LRF (a field expansion) does not, so compare without it
        scale = np.abs(d).sum() / d.size
        q2 = synth.water_box(32, 20261018)
        q2.use_LRF = 0
        g2 = Qnb(q2)
        g2.make_pair_lists(q2.xtop, **cuts)
        d2, E2, _ = g2.pot_energy_nonbonds(q2.xtop, lam)
        assert np.abs(d2.sum(axis=0)).max() <= 1e-6 * scale * d2.shape[0] ** 0.5
        assert_energy("ww.el", E2[4], E[4])
        assert_energy("ww.vdw", E2[5], E[5])
        xs = q2.xtop + q2.boxlength[None, :] * np.array([1.0, -2.0, 3.0])[None, :]
        g2.make_pair_lists(xs, **cuts)
        d3, E3, _ = g2.pot_energy_nonbonds(xs, lam)
        assert_energy("ww.el shifted", E3[4], E2[4], tol=1e-6)
        assert_energy("ww.vdw shifted", E3[5], E2[5], tol=1e-6)
        g2.close()
    finally:
        g.close()


@pytest.mark.parametrize("case", [c for c in small_systems() if c[0] in ("box_water", "box_solute_q", "box_solute_rowimage")],
                         ids=lambda c: c[0])
def test_mc_volume_save_restore(case):
    """MC_volume (md.f90:1976-2284): lists + LRF of the old box are kept while a scaled box is tried, and come back
    unchanged when the move is rejected."""
    from oracle.pyoracle import Oracle
    from q6_b200.engine import Qnb
    name, q, cuts, lam = case
    lam = np.array(lam)
    g, o = Qnb(q), Oracle(q)
    try:
        x0 = q.xtop
        rng = np.random.default_rng(3)
        x1 = x0 + rng.normal(0, 0.02, x0.shape)          # some steps after the list build
        g.make_pair_lists(x0, **cuts)
        o.make_pair_lists(x0, **cuts)
        lists0 = [sorted_pairs(g.export_list(w, 1)[0]) for w in range(5)]
        lrf0 = g.export_lrf()
        d0, E0, EQ0 = g.pot_energy_nonbonds(x1, lam)
        g.save_lists()
        # trial move: box and coordinates scaled by 1 %, new lists, new energy (md.f90:2087-2177)
        box0 = np.array(q.boxlength, dtype=float)
        s = 1.01
        g.update_box(box0 * s)
        g.make_pair_lists(x1 * s, **cuts)
        dt, Et, EQt = g.pot_energy_nonbonds(x1 * s, lam)
        assert abs(Et[4] - E0[4]) > 0 or abs(Et[0] - E0[0]) > 0      # it really was another state
        # rejected: everything back
        g.restore_lists()
        for w in range(5):
            assert np.array_equal(sorted_pairs(g.export_list(w, 1)[0]), lists0[w]), f"{LIST_NAMES[w]} list differs after restore"
        assert np.array_equal(g.export_lrf(), lrf0), "LRF moments are not restored bit for bit"
        d2, E2, EQ2 = g.pot_energy_nonbonds(x1, lam)
        # same lists as sets; the order inside a row (hence FP32 summation order) may differ between two builds
        assert rel_rms(d2, d0) < 2e-6
        assert np.allclose(E2, E0, rtol=1e-10, atol=1e-9) and np.allclose(EQ2, EQ0, rtol=1e-10, atol=1e-9)
        # and it is still the reference's answer for the old box
        do, Eo, EQo = o.pot_energy_nonbonds(x1, lam)
        assert rel_rms(d2, do) <= FORCE_REL_RMS
        for k, nm in enumerate(("pp.el", "pp.vdw", "pw.el", "pw.vdw", "ww.el", "ww.vdw", "LRF")):
            assert_energy(nm, E2[k], Eo[k])
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("case", [c for c in small_systems() if c[0] in ("sph_evb2", "box_solute_q", "sph_anyatom")],
                         ids=lambda c: c[0])
def test_qcp_beads(case):
    """qcp_run (qcp.f90:319-372): per-bead pot_energy(E,EQ,.false.) with the path-integral atoms displaced; the batched
    entry must give, bead by bead, what the reference's loop gives."""
    from oracle.pyoracle import Oracle
    from q6_b200.engine import Qnb
    name, q, cuts, lam = case
    lam = np.array(lam)
    g, o = Qnb(q), Oracle(q)
    try:
        x0 = q.xtop
        g.make_pair_lists(x0, **cuts)
        o.make_pair_lists(x0, **cuts)
        rng = np.random.default_rng(9)
        qat = np.asarray(q.iqseq[:q.nqat])
        atoms = qat[rng.choice(q.nqat, size=min(3, q.nqat), replace=False)]     # the qcp_atom selection
        nbeads = 8
        coord = rng.normal(0, 0.08, (nbeads, atoms.size, 3))
        coord -= coord.mean(axis=0, keepdims=True)                              # qcp_center
        EQ = g.qcp_beads(x0, atoms, coord, lam)
        for b in range(nbeads):
            xb = x0.copy()
            xb[atoms - 1] += coord[b]
            _, _, EQo = o.pot_energy_nonbonds(xb, lam, md=False)
            for s in range(q.nstates):
                for k, nm in enumerate(("qq.el", "qq.vdw", "qp.el", "qp.vdw", "qw.el", "qw.vdw")):
                    assert_energy(f"bead {b} state {s + 1} {nm}", EQ[b, s, k], EQo[s, k])
        # the beads really differ from each other
        assert np.abs(EQ - EQ[0]).max() > 0
        # and the handle is unharmed: a normal step afterwards still matches
        _check_step(g, o, q, x0, lam)
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("case", [c for c in small_systems() if c[0] in ("box_water_rowimage", "box_solute_rowimage", "box_water")],
                         ids=lambda c: c[0])
def test_unwrapped_periodic_coordinates(case):
    """Qdyn6 does not keep molecules inside the box between put_back_in_box calls: whole charge groups shifted by
    multiples of the box length must give the same lists, and the same energies / gradient as the oracle."""
    from oracle.pyoracle import Oracle
    from q6_b200.engine import Qnb
    name, q, cuts, lam = case
    lam = np.array(lam)
    rng = np.random.default_rng(17)
    x = q.xtop.copy()
    box = np.asarray(q.boxlength, dtype=float)
    cgp = np.asarray(q.cgp).reshape(-1, 3)          # (iswitch, first, last), 1-based
    cgpatom = np.asarray(q.cgpatom)
    for g_ in range(q.ncgp):
        sh = rng.integers(-2, 3, size=3) * box
        atoms = cgpatom[cgp[g_, 1] - 1:cgp[g_, 2]] - 1
        x[atoms] += sh
    g, o = Qnb(q), Oracle(q)
    try:
        cg = g.make_pair_lists(x, **cuts)
        co = o.make_pair_lists(x, **cuts)
        assert np.array_equal(cg[:5], co[:5])
        _check_lists(g, o, q.nstates)
        if q.use_LRF:
            lg, lo = g.export_lrf(), o.export_lrf()
            scale = np.abs(lo).max(axis=0) + 1e-300
            tol = np.where(np.arange(43) < 16, 1e-9, 2e-5)
            assert np.all(np.abs(lg - lo) <= tol * scale + 1e-12), "LRF moments differ"
        _check_step(g, o, q, x, lam)
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("case", [c for c in small_systems() if c[0] in ("sph_fep2", "sph_excl_shell", "sph_noq")],
                         ids=lambda c: c[0])
def test_solvent_restraints(case):
    """restrain_solvent + watpol (nonbondene.f90:6466-6746) inside the device step: gradient and energies of the step
    with QNB_FLAG_SOLVENT_RESTRAINTS = nonbonded oracle + restraint oracle; shell populations exact."""
    from oracle import pyoracle
    from oracle.pyoracle import Oracle
    from q6_b200.engine import Qnb, wat_shells
    name, q, cuts, lam = case
    lam = np.array(lam)
    rwat = float(np.linalg.norm(q.xtop[q.nat_solute:] - np.asarray(q.xpcent), axis=1).max()) - 0.3
    par = wat_shells(q.xpcent, rwat, crgQtot=-1.0)
    tc = np.array([0.02, -0.01, 0.005][:par.nwpolr_shell])
    g, o = Qnb(q), Oracle(q)
    try:
        rng = np.random.default_rng(4)
        x = q.xtop + rng.normal(0, 0.03, q.xtop.shape)
        g.make_pair_lists(x, **cuts)
        o.make_pair_lists(x, **cuts)
        g.set_solvent_restraints(par, tc)
        dg, Eg, EQg = g.pot_energy_nonbonds(x, lam, restraints=True)
        Er, ts, ns = g.last_restraints()
        do, Eo, EQo = o.pot_energy_nonbonds(x, lam)
        dr, Ero, tso, nso = pyoracle.solvent_restraints(q, par, tc, x)
        assert ns.sum() > 0 and np.array_equal(ns, nso), (ns, nso)
        assert rel_rms(dg, do + dr) <= FORCE_REL_RMS
        # the restraint part on its own (FP64 on both sides)
        d0, _, _ = g.pot_energy_nonbonds(x, lam)
        assert rel_rms(dg - d0, dr) <= 1e-6
        assert_energy("solvent_radial", Er[0], Ero[0])
        assert_energy("water_pol", Er[1], Ero[1])
        assert np.allclose(ts, tso, rtol=1e-12)
        for k, nm in enumerate(("pp.el", "pp.vdw", "pw.el", "pw.vdw", "ww.el", "ww.vdw", "LRF")):
            assert_energy(nm, Eg[k], Eo[k])
        # new theta_corr (every itdis steps on the host) reaches the captured step
        tc2 = tc + 0.1
        g.set_theta_corr(tc2)
        dg2, _, _ = g.pot_energy_nonbonds(x, lam, restraints=True)
        dr2, Ero2, _, _ = pyoracle.solvent_restraints(q, par, tc2, x)
        assert rel_rms(dg2 - d0, dr2) <= 1e-6
        assert_energy("water_pol (new theta_corr)", g.last_restraints()[0][1], Ero2[1])
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("one_graph", [False, True, "parent"], ids=["graph_per_window", "one_graph", "parent_graph"])
def test_batched_windows_equal_single_calls(one_graph, monkeypatch):
    """qnb_build_lists_batch / qnb_nonbond_batch over W lambda windows of one FEP system (own coordinates and lambda per
    window, run_excl_test.sh:92-125) == W single-system calls: same list counts, gradients to FP64 summation-order noise,
    energies likewise; every window also checked against the oracle."""
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    from q6_b200.engine import Qnb, QnbBatch
    if one_graph == "parent":
        monkeypatch.setenv("QNB_BATCH_PARENT", "1")     # the windows' step graphs as children of one parent graph
    elif one_graph:
        monkeypatch.setenv("QNB_BATCH_GRAPH", "1")      # one launch per kernel type for all windows (k_batched)
    q = synth.solvated_sphere(14.0, 0.0, 14, 2, 97, fep="annihilate")
    cuts = common.sph_cuts(8.0)
    W = 5
    rng = np.random.default_rng(5)
    xs = [q.xtop + rng.normal(0.0, 0.02, q.xtop.shape) for _ in range(W)]
    lams = [np.array([1.0 - 0.2 * k, 0.2 * k]) for k in range(W)]
    hs = [Qnb(q) for _ in range(W)]
    b = QnbBatch(hs)
    cb = b.make_pair_lists(xs, **cuts, counts=True)
    d, E, EQ = b.pot_energy_nonbonds(xs, lams)
    d = [v.copy() for v in d]; E = [v.copy() for v in E]; EQ = [v.copy() for v in EQ]
    # a second batched step on the same inputs adds to d unless the caller says d is zero (pot_energy adds to d)
    d2, _, _ = b.pot_energy_nonbonds(zero_d=False)
    for k in range(W):
        assert np.allclose(d2[k], 2.0 * d[k], rtol=1e-9, atol=1e-9)
    d3, _, _ = b.pot_energy_nonbonds()          # QNB_FLAG_D_IS_ZERO: written, whatever the arrays held
    for k in range(W):
        assert np.allclose(d3[k], d[k], rtol=1e-9, atol=1e-9)
    o = Oracle(q)
    for k in range(W):
        g = Qnb(q)
        c1 = g.make_pair_lists(xs[k], **cuts)
        d1, E1, EQ1 = g.pot_energy_nonbonds(xs[k], lams[k])
        assert np.array_equal(cb[k][:5], c1[:5])
        assert common.rel_rms(d[k], d1) <= 1e-12
        assert np.allclose(E[k], E1, rtol=1e-11, atol=1e-9) and np.allclose(EQ[k], EQ1, rtol=1e-11, atol=1e-9)
        co = o.make_pair_lists(xs[k], **cuts)
        do, Eo, EQo = o.pot_energy_nonbonds(xs[k], lams[k])
        assert np.array_equal(c1[:5], co[:5])
        assert common.rel_rms(d[k], do) <= common.FORCE_REL_RMS
        common.assert_energy("E", E[k], Eo)
        common.assert_energy("EQ", EQ[k], EQo)
        g.close()
    # errors: a handle given twice is refused
    with pytest.raises(Exception):
        QnbBatch([hs[0], hs[0]])
    for g in hs:
        g.close()


def test_registered_host_buffers_same_results():
    """qnb_register_host_buffers (x uploaded without staging, gradient added into d by the device) gives what the staged
    path gives: d is ADDED to (potene.f90:109 zeroes it, bonded terms may already be in it), other arrays still work."""
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    q = synth.solvated_sphere(14.0, 7.0, 10, 2, 41, fep="evb")
    cuts, lam = common.sph_cuts(8.0), np.array([0.3, 0.7])
    g = Qnb(q)
    x = q.xtop.copy()
    g.make_pair_lists(x, **cuts)
    d0, E0, EQ0 = g.pot_energy_nonbonds(x, lam)
    d0 = d0.copy()
    d = np.full((q.natom, 3), 0.25)            # pre-loaded, e.g. with bonded terms
    g.register_host_buffers(x.reshape(-1), d.reshape(-1))
    _, E1, EQ1 = g.pot_energy_nonbonds(x, lam, d=d)
    assert rel_rms(d - 0.25, d0) <= 1e-12 and np.allclose(E1, E0, rtol=1e-11, atol=1e-9) and np.allclose(EQ1, EQ0, rtol=1e-11, atol=1e-9)
    x[:] = x + 0.01                            # the registered array is read anew every step
    d[:] = 0.0
    g.pot_energy_nonbonds(x, lam, d=d)
    d2, _, _ = g.pot_energy_nonbonds(x.copy(), lam)      # unregistered arrays: the staged path
    assert rel_rms(d, d2) <= 1e-12
    g.release_host_buffers()
    d[:] = 0.0
    g.pot_energy_nonbonds(x, lam, d=d)
    assert rel_rms(d, d2) <= 1e-12
    g.close()


def test_bound_step_equals_the_plain_call():
    """Qnb.bind_step (pointers held once) == pot_energy_nonbonds on the same arrays, and follows changes of x and lambda."""
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    q = synth.solvated_sphere(14.0, 7.0, 10, 2, 43, fep="evb")
    cuts = common.sph_cuts(8.0)
    g = Qnb(q)
    x = q.xtop.copy()
    lam = np.array([0.6, 0.4])
    d = np.zeros((q.natom, 3))
    g.make_pair_lists(x, **cuts)
    step, E, EQ = g.bind_step(x.reshape(-1), lam, d.reshape(-1))
    step()
    d1, E1, EQ1 = g.pot_energy_nonbonds(x, lam)
    assert rel_rms(d, d1) <= 1e-12 and np.allclose(E, E1, rtol=1e-11, atol=1e-9) and np.allclose(EQ, EQ1, rtol=1e-11, atol=1e-9)
    x += 0.01; lam[:] = (0.2, 0.8); d[:] = 0.0
    step()
    d2, E2, EQ2 = g.pot_energy_nonbonds(x, lam)
    assert rel_rms(d, d2) <= 1e-12 and np.allclose(E, E2, rtol=1e-11, atol=1e-9) and np.allclose(EQ, EQ2, rtol=1e-11, atol=1e-9)
    g.close()
