"""The compiled (C++) host side above the C ABI -- q6_b200/host -- against the Python mirror and the oracle.

The reference host is Fortran; with no Fortran compiler in the image its role is played by qdyn_host.cpp (topo_read,
qatom_load_fep, prep_sim, make_qconn, initial SHAKE, Nonbonded::make_pair_lists / pot_energy_nonbonds, write_out).
CPU tests: the tables it prepares from the reference's shipped topology / FEP files equal the Python readers' tables
field by field, the oracle fed with them reproduces the reference's step-0 goldens, and its write_out prints the very
line eval_test.sh:99-101 greps for.  GPU test: the Nonbonded class through libqnb against the committed goldens.
"""
import ctypes as C
import os

import numpy as np
import pytest

import common
from common import golden_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "q6_b200", "host", "libqdynhost.so")
REF = "/root/reference/tests"
have_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

FILES = {
    "c1_sph": (f"{REF}/basic_tests/prep_SPH/lig_w.top", f"{REF}/basic_tests/prep_SPH/lig_w.fep"),
    "c1_pbc": (f"{REF}/basic_tests/prep_PBC/lig_w.top", f"{REF}/basic_tests/prep_PBC/lig_w.fep"),
    "c4_evb": (f"{REF}/exclude_tests/inputs/2cjpFH_ionres_oplsa.top", f"{REF}/exclude_tests/inputs/lig.fep"),
}


def _ensure_built():
    """libqdynhost.so / qdyn_nb are built by __graft_entry__.build(); build them here if a fresh checkout runs the tests."""
    import subprocess
    host = os.path.join(ROOT, "q6_b200", "host")
    if not (os.path.exists(LIB) and os.path.exists(os.path.join(host, "qdyn_nb"))):
        subprocess.check_call(["make", "-s", "-C", host])


def host_lib():
    from q6_b200 import engine
    from q6_b200.system import qnb_system
    engine.load_library()          # libqdynhost.so links against libqnb.so
    _ensure_built()
    lib = C.CDLL(LIB)
    H = C.c_void_p
    PD, PL = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib.qhost_last_error.restype = C.c_char_p
    lib.qhost_open.restype = C.c_int
    lib.qhost_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(H)]
    lib.qhost_from_system.restype = C.c_int
    lib.qhost_from_system.argtypes = [C.POINTER(qnb_system), PD, C.POINTER(H)]
    lib.qhost_system.restype = C.POINTER(qnb_system)
    lib.qhost_system.argtypes = [H]
    lib.qhost_xtop.restype = PD
    lib.qhost_xtop.argtypes = [H]
    lib.qhost_boxlength.restype = PD
    lib.qhost_boxlength.argtypes = [H]
    lib.qhost_shard.argtypes = [H, C.c_int, C.c_int]
    lib.qhost_constraint_count.restype = C.c_int64
    lib.qhost_constraint_count.argtypes = [H]
    lib.qhost_constraint_molecules.restype = C.c_int64
    lib.qhost_constraint_molecules.argtypes = [H, C.POINTER(C.c_int32), C.c_int64]
    lib.qhost_initial_constraint.argtypes = [H, PD, C.POINTER(C.c_int)]
    lib.qhost_attach_gpu.argtypes = [H, C.c_int]
    lib.qhost_make_pair_lists.argtypes = [H, PD] + [C.c_double] * 7 + [PL]
    lib.qhost_pot_energy_nonbonds.argtypes = [H, PD, PD, C.c_int, PD, PD, PD]
    lib.qhost_write_out.restype = C.c_char_p
    lib.qhost_write_out.argtypes = [H, PD, PD, PD, C.c_int]
    lib.qhost_close.argtypes = [H]
    return lib


def _open(lib, name, use_lrf=1):
    top, fep = FILES[name]
    h = C.c_void_p()
    rc = lib.qhost_open(top.encode(), fep.encode(), use_lrf, -1, C.byref(h))
    assert rc == 0, lib.qhost_last_error().decode()
    return h


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _struct_arrays(st):
    """Every array of a qnb_system as numpy copies, sized from its scalars."""
    nat, nsol, nq, ns = st.natom, st.nat_solute, st.nqat, st.nstates
    ncgpatom = max([st.cgp[3 * g + 2] for g in range(st.ncgp)] + [0])     # cgp(ncgp)%last
    sizes = dict(cgp=3 * st.ncgp, cgpatom=ncgpatom, excl=nat, iqatom=nat, iqseq=nq, iac=nat, crg=nat, iaclib=7 * st.natyps,
                 ljcod=st.num_atyp ** 2, listex=st.max_nbr_range * nsol, list14=st.max_nbr_range * nsol,
                 listexlong=2 * st.nexlong, list14long=2 * st.n14long, qcrg=nq * ns, qiac=nq * ns, qavdw=3 * st.nqlib,
                 qbvdw=3 * st.nqlib, sc_lookup=nq * (st.natyps + nq) * ns, iqexpnb=st.nqexpnb, jqexpnb=st.nqexpnb,
                 el_scale_iq=st.nel_scale, el_scale_jq=st.nel_scale, el_scale=st.nel_scale * ns, qconn=ns * nsol * nq)
    return {k: np.array(getattr(st, k)[:n]) for k, n in sizes.items()}


def _scalars():
    from q6_b200.system import qnb_system
    return [n for n, t in qnb_system._fields_ if t in (C.c_int32, C.c_double)]


@have_ref
@pytest.mark.parametrize("name", list(FILES))
def test_cpp_host_tables_equal_python_mirror(name):
    """topo_read + qatom_load_fep + prep_sim in C++ == q6_b200.topo/fep/system on the reference's shipped inputs:
    every scalar and every array of qnb_system, bit for bit."""
    from q6_b200.fep import load_fep
    from q6_b200.system import build_system
    from q6_b200.topo import topo_read
    lib = host_lib()
    h = _open(lib, name)
    try:
        cs = lib.qhost_system(h).contents
        t = topo_read(FILES[name][0])
        py, keep = build_system(t, load_fep(FILES[name][1], t), use_LRF=True).as_struct()
        for n in _scalars():
            assert getattr(cs, n) == getattr(py, n), n
        assert list(cs.xpcent) == list(py.xpcent)
        a, b = _struct_arrays(cs), _struct_arrays(py)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        x = np.ctypeslib.as_array(lib.qhost_xtop(h), shape=(cs.natom, 3))
        assert np.array_equal(x, t.xtop)
        if t.use_PBC:
            assert np.array_equal(np.ctypeslib.as_array(lib.qhost_boxlength(h), shape=(3,)), t.boxlength)
    finally:
        lib.qhost_close(h)


@have_ref
@pytest.mark.parametrize("name", ["c1_sph", "c1_pbc"])
def test_cpp_host_reproduces_reference_step0(name):
    """The compiled host end to end on the CPU side: shipped files -> tables -> initial SHAKE (bit-identical to the
    fixture's post-SHAKE coordinates) -> the oracle on those tables gives the reference's step-0 energies, and
    write_out prints the line eval_test.sh compares."""
    from oracle import pyoracle
    from test_oracle_golden import STEP0_GOLDEN, step0_terms
    lib = host_lib()
    h = _open(lib, name)
    try:
        cs = lib.qhost_system(h)
        st = cs.contents
        assert lib.qhost_constraint_count(h) == 3 * st.nwat
        q, cuts, lam, z = golden_system(name)
        x = np.ctypeslib.as_array(lib.qhost_xtop(h), shape=(st.natom, 3)).copy()
        nit = C.c_int()
        assert lib.qhost_initial_constraint(h, _dp(x), C.byref(nit)) == 0, lib.qhost_last_error().decode()
        assert np.array_equal(x, z["x_step0"])
        # the oracle driven by the C++ host's struct
        ol = pyoracle.load()
        oh = ol.qo_create(cs)
        assert oh, ol.qo_last_error().decode()
        if st.use_PBC:
            b = np.ctypeslib.as_array(lib.qhost_boxlength(h), shape=(3,)).copy()
            ib = 1.0 / b
            ol.qo_update_box(oh, _dp(b), _dp(ib))
        counts = np.zeros(8, np.int64)
        c7 = [cuts[k] for k in common.CUT_KEYS]
        assert ol.qo_make_pair_lists(oh, _dp(x.reshape(-1)), *c7, counts.ctypes.data_as(C.POINTER(C.c_int64))) == 0
        d, E, EQ = np.zeros(3 * st.natom), np.zeros(7), np.zeros(6 * st.nstates)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        ol.qo_nonbond(oh, _dp(x.reshape(-1)), _dp(lam), 3, _dp(d), _dp(E), _dp(EQ))
        ol.qo_destroy(oh)
        assert np.array_equal(counts, z["counts_step0"])
        assert np.allclose(E, z["E_step0"], rtol=1e-12, atol=1e-9)
        got = step0_terms(E, EQ.reshape(-1, 6))
        for v, want in zip(got, STEP0_GOLDEN[name]):
            assert abs(v - want) <= 0.005 + 1e-9
        text = lib.qhost_write_out(h, _dp(E), _dp(EQ), _dp(lam), 0).decode()
        want = {"c1_sph": "Q-surr. 1 1.0000      3.12    139.43", "c1_pbc": "Q-surr. 1 1.0000    -31.30    228.64"}[name]
        assert want in text.split("\n"), text          # INIT_QSURR_BM, eval_test.sh:99-101
        assert "at step      0" in text
    finally:
        lib.qhost_close(h)


def test_cpp_host_from_struct_and_sharding():
    """system_from_struct deep-copies tables prepared elsewhere; shard() is distribute_nonbonds with equal shares."""
    from q6_b200.system import distribute_nonbonds
    lib = host_lib()
    q, cuts, lam, z = golden_system("c4_evb")
    st, keep = q.as_struct()
    h = C.c_void_p()
    assert lib.qhost_from_system(C.byref(st), None, C.byref(h)) == 0, lib.qhost_last_error().decode()
    try:
        cs = lib.qhost_system(h).contents
        a, b = _struct_arrays(cs), _struct_arrays(st)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert lib.qhost_shard(h, 1, 4) == 0
        cs = lib.qhost_system(h).contents
        assert (cs.pp_start, cs.pp_end) == distribute_nonbonds(np.ones(q.ncgp_solute), 4)[1]
        assert (cs.ww_start, cs.ww_end) == distribute_nonbonds(np.ones(q.nwat), 4)[1]
        assert cs.is_master == 0
    finally:
        lib.qhost_close(h)


def test_cpp_host_dies_without_gpu_and_on_bad_files(tmp_path):
    """Error behaviour: topo_read's message for an unreadable topology; Nonbonded refuses to exist without a GPU."""
    from q6_b200 import engine
    lib = host_lib()
    bad = tmp_path / "bad.top"
    bad.write_text("not a topology\n1 2 x\n")
    h = C.c_void_p()
    assert lib.qhost_open(str(bad).encode(), b"", 1, -1, C.byref(h)) != 0
    assert "Could not read topology file" in lib.qhost_last_error().decode()
    assert lib.qhost_open(b"/nonexistent.top", b"", 1, -1, C.byref(h)) != 0
    if engine.load_library().qnb_device_count() > 0:
        return
    q, cuts, lam, z = golden_system("c1_sph")
    st, keep = q.as_struct()
    assert lib.qhost_from_system(C.byref(st), None, C.byref(h)) == 0
    try:
        assert lib.qhost_attach_gpu(h, 0) != 0
        assert "qnb_init" in lib.qhost_last_error().decode()
    finally:
        lib.qhost_close(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_sph", "c1_pbc", "c4_evb"])
def test_cpp_nonbonded_class_on_gpu(name):
    """qdyn::Nonbonded (make_pair_lists / pot_energy_nonbonds with the reference's argument lists) through libqnb on
    the reference's shipped systems: list sizes, energies and gradient against the committed oracle results, and for
    the two basic_tests systems the reference's own step-0 numbers."""
    from test_oracle_golden import STEP0_GOLDEN, step0_terms
    lib = host_lib()
    q, cuts, lam, z = golden_system(name)
    st, keep = q.as_struct()
    h = C.c_void_p()
    box = np.ascontiguousarray(q.boxlength, dtype=np.float64)
    assert lib.qhost_from_system(C.byref(st), _dp(box), C.byref(h)) == 0, lib.qhost_last_error().decode()
    try:
        assert lib.qhost_attach_gpu(h, 0) == 0, lib.qhost_last_error().decode()
        cases = [("", q.xtop)] + ([("_step0", z["x_step0"])] if name in STEP0_GOLDEN else [])
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        for tag, x in cases:
            x = np.ascontiguousarray(x, dtype=np.float64)
            counts = np.zeros(8, np.int64)
            c7 = [cuts[k] for k in common.CUT_KEYS]
            rc = lib.qhost_make_pair_lists(h, _dp(x), *c7, counts.ctypes.data_as(C.POINTER(C.c_int64)))
            assert rc == 0, lib.qhost_last_error().decode()
            assert np.array_equal(counts[:5], z["counts" + tag][:5])
            d, E, EQ = np.zeros((q.natom, 3)), np.zeros(7), np.zeros(6 * q.nstates)
            rc = lib.qhost_pot_energy_nonbonds(h, _dp(x), _dp(lam), 1, _dp(d), _dp(E), _dp(EQ))
            assert rc == 0, lib.qhost_last_error().decode()
            assert common.rel_rms(d, z["d" + tag]) <= common.FORCE_REL_RMS
            for k in range(7):
                common.assert_energy(f"E[{k}]", E[k], z["E" + tag][k])
            common.assert_energy("EQ", EQ.reshape(-1, 6), z["EQ" + tag])
            if tag:
                for v, want in zip(step0_terms(E, EQ.reshape(-1, 6)), STEP0_GOLDEN[name]):
                    assert abs(v - want) <= 0.005 + 1e-9
    finally:
        lib.qhost_close(h)


# ---------------------------------------------------------------------------------------------------------------------
# Synthetic systems written as Qdyn6 input files (q6_b200.synth.write_files): these run where /root/reference does not
# exist (the GPU box), so both readers and the compiled driver are covered there too.

def _small(names=None):
    return [c for c in common.small_systems() + common.variant_systems() if names is None or c[0] in names]


@pytest.mark.parametrize("case", _small(), ids=lambda c: c[0])
def test_written_input_files_round_trip_through_both_readers(case, tmp_path):
    """write_files -> (Python topo_read/load_fep/build_system, C++ topo_read/qatom_load_fep/prep_sim) reproduces every
    table of the synthetic system bit for bit: sphere and box, 0-2 states, softcore, soft pairs, per-state Q-bonds,
    long-range exclusion / 1-4 lists, excluded shells, any-atom cut-offs, no solute, no water."""
    from q6_b200 import synth
    from q6_b200.fep import load_fep
    from q6_b200.system import QSystem, build_system
    from q6_b200.topo import topo_read
    name, q, cuts, lam = case
    top, fep = str(tmp_path / "s.top"), str(tmp_path / "s.fep")
    synth.write_files(q, top, fep)
    t = topo_read(top)
    r = build_system(t, load_fep(fep, t) if q.nqat else None, use_LRF=bool(q.use_LRF))
    for k in QSystem._SCALARS:
        assert getattr(q, k) == getattr(r, k), k
    for k in QSystem._ARRAYS:
        a, b = getattr(q, k), getattr(r, k)
        a = np.zeros(0) if a is None else np.asarray(a).reshape(-1)
        b = np.zeros(0) if b is None else np.asarray(b).reshape(-1)
        assert a.size == b.size and np.array_equal(a, b), k
    lib = host_lib()
    h = C.c_void_p()
    rc = lib.qhost_open(top.encode(), fep.encode() if q.nqat else b"", int(q.use_LRF), -1, C.byref(h))
    assert rc == 0, lib.qhost_last_error().decode()
    try:
        cs = lib.qhost_system(h).contents
        py, keep = q.as_struct()
        for n in _scalars():
            assert getattr(cs, n) == getattr(py, n), n
        assert list(cs.xpcent) == list(py.xpcent)
        a, b = _struct_arrays(cs), _struct_arrays(py)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(np.ctypeslib.as_array(lib.qhost_xtop(h), shape=(q.natom, 3)), q.xtop)
        # initial SHAKE of the jittered rigid waters: the C++ host and the Python restatement agree bit for bit
        if q.nwat:
            from oracle import pyoracle
            assert lib.qhost_constraint_count(h) == 3 * q.nwat
            first = np.zeros(q.nwat + 1, np.int32)            # const_mol(:): three constraints per water, molecule by molecule
            assert lib.qhost_constraint_molecules(h, first.ctypes.data_as(C.POINTER(C.c_int32)), len(first)) == q.nwat + 1
            assert np.array_equal(first, 3 * np.arange(q.nwat + 1))
            x = q.xtop.copy()
            assert lib.qhost_initial_constraint(h, _dp(x), None) == 0, lib.qhost_last_error().decode()
            xs, nits = pyoracle.initial_constraint_x(t)
            assert np.array_equal(x, xs)
            assert np.abs(x - q.xtop).max() > 1e-4          # the jitter really broke the constraints
    finally:
        lib.qhost_close(h)


def _driver_rows(text):
    """The numbers of the energy summary rows qdyn_nb prints (write_out formats)."""
    rows = {}
    for ln in text.splitlines():
        for key in ("solute-solvent", "solute", "solvent", "LRF"):
            if ln.startswith(key.ljust(16)):
                rows[key] = [float(v) for v in ln[16:].split()]
                break
        for key in ("Q-Q", "Q-prot", "Q-wat", "Q-surr."):
            if ln.startswith(key.ljust(7)) and len(ln) > 16 and ln[7:9].strip().isdigit():
                rows.setdefault(key, {})[int(ln[7:9])] = [float(ln[16:26]), float(ln[26:36])]
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("case", _small(("sph_evb2", "box_solute_q")), ids=lambda c: c[0])
def test_compiled_driver_end_to_end_on_gpu(case, tmp_path):
    """q6_b200/host/qdyn_nb on written input files: read -> prepare -> initial SHAKE -> make_pair_lists -> pot_energy ->
    write_out, everything above the C ABI in C++; the printed step-0 summary against the oracle at the same
    (post-SHAKE) coordinates, to the two printed decimals."""
    import subprocess
    from oracle import pyoracle
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    from q6_b200.topo import topo_read
    name, q, cuts, lam = case
    top, fep = str(tmp_path / "s.top"), str(tmp_path / "s.fep")
    synth.write_files(q, top, fep)
    rc2 = float(np.sqrt(cuts["Rcpp2"]))
    cmd = [os.path.join(ROOT, "q6_b200", "host", "qdyn_nb"), top, fep, "--q_atom", repr(cuts["Rq"]), "--lrf",
           repr(cuts["RcLRF"]), "--solute_solute", repr(rc2), "--solute_solvent", repr(float(np.sqrt(cuts["Rcpw2"]))),
           "--solvent_solvent", repr(float(np.sqrt(cuts["Rcww2"]))), "--lambda", ",".join(repr(float(v)) for v in lam),
           "--steps", "30"]
    if not q.use_LRF:
        cmd.append("--no-lrf")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = _driver_rows(out.stdout)
    xs, _ = pyoracle.initial_constraint_x(topo_read(top))
    o = Oracle(q)
    c = o.make_pair_lists(xs, **cuts)
    d, E, EQ = o.pot_energy_nonbonds(xs, lam)
    o.close()
    counts = [int(v) for v in out.stdout.splitlines()[[i for i, l in enumerate(out.stdout.splitlines())
                                                       if l.startswith("solute-solute")][0] + 1].split()]
    assert counts == [int(v) for v in c[:5]]
    tol = lambda v: 0.0051 + 1e-6 * abs(v)
    assert abs(rows["solute"][0] - E[0]) <= tol(E[0]) and abs(rows["solute"][1] - E[1]) <= tol(E[1])
    assert abs(rows["solvent"][0] - E[4]) <= tol(E[4]) and abs(rows["solvent"][1] - E[5]) <= tol(E[5])
    assert abs(rows["solute-solvent"][0] - E[2]) <= tol(E[2]) and abs(rows["solute-solvent"][1] - E[3]) <= tol(E[3])
    if q.use_LRF:
        assert abs(rows["LRF"][0] - E[6]) <= tol(E[6])
    for s in range(q.nstates):
        for key, k0 in (("Q-Q", 0), ("Q-prot", 2), ("Q-wat", 4)):
            got = rows[key][s + 1]
            assert abs(got[0] - EQ[s, k0]) <= tol(EQ[s, k0]) and abs(got[1] - EQ[s, k0 + 1]) <= tol(EQ[s, k0 + 1]), key
    assert "ms per step" in out.stdout


@pytest.mark.parametrize("name", ["sph_evb2", "sph_arith_evb"])
def test_softcore_max_potential_tables_agree_between_readers(name, tmp_path):
    """[FEP] softcore_use_max_potential on (qatom.f90:1932-1990): sc_lookup is derived from alpha and the LJ parameters,
    with different formulas for the geometric and the arithmetic rule.  No shipped input uses it; the C++ and the Python
    reader must at least agree with each other bit for bit, and differ from the plain-alpha tables."""
    from q6_b200 import synth
    from q6_b200.fep import load_fep
    from q6_b200.system import build_system
    from q6_b200.topo import topo_read
    q = [c for c in _small((name,))][0][1]
    top, fep = str(tmp_path / "s.top"), str(tmp_path / "s.fep")
    synth.write_files(q, top, fep)
    text = open(fep).read().replace("[FEP]\n", "[FEP]\nsoftcore_use_max_potential on\n", 1)
    open(fep, "w").write(text)
    t = topo_read(top)
    py, keep = build_system(t, load_fep(fep, t), use_LRF=True).as_struct()
    lib = host_lib()
    h = C.c_void_p()
    assert lib.qhost_open(top.encode(), fep.encode(), 1, -1, C.byref(h)) == 0, lib.qhost_last_error().decode()
    try:
        cs = lib.qhost_system(h).contents
        a, b = _struct_arrays(cs), _struct_arrays(py)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        plain, _ = q.as_struct()
        assert not np.array_equal(a["sc_lookup"], _struct_arrays(plain)["sc_lookup"])
        assert np.isfinite(a["sc_lookup"]).all() and (a["sc_lookup"] >= 0).all()
    finally:
        lib.qhost_close(h)


@pytest.mark.parametrize("key", ["offset", "offset_residue", "offset_name"])
def test_fep_offsets_reproduce_the_same_tables(key, tmp_path):
    """[FEP] offset / offset_residue / offset_name (qatom.f90:383-432): topology numbers in [atoms], [change_bonds],
    [excluded_pairs] and [PBC] switching_atom are relative to the offset.  A file rewritten with an offset must give the
    very tables of the original one, in both readers."""
    from q6_b200 import synth
    from q6_b200.fep import load_fep
    from q6_b200.system import build_system
    from q6_b200.topo import topo_read
    q = [c for c in _small(("box_solute_q",))][0][1]          # periodic (switching atom), EVB (per-state Q-bonds)
    top, fep0, fep1 = str(tmp_path / "s.top"), str(tmp_path / "s0.fep"), str(tmp_path / "s1.fep")
    synth.write_files(q, top, fep0)
    t = topo_read(top)
    # the synthetic solute is one residue starting at atom 1: offset_residue 1 / offset_name SYN give offset 0
    off = 3 if key == "offset" else 0
    head = {"offset": f"offset {off}", "offset_residue": "offset_residue 1", "offset_name": "offset_name SYN"}[key]
    out, sec = [], None
    for ln in open(fep0).read().splitlines():
        if ln.startswith("["):
            sec = ln.strip("[]").lower()
            out.append(ln)
            if sec == "fep":
                out.append(head)
            continue
        tok = ln.split()
        if sec == "atoms":
            tok[1] = str(int(tok[1]) - off)
        elif sec == "change_bonds":
            tok[0], tok[1] = str(int(tok[0]) - off), str(int(tok[1]) - off)
        elif sec == "pbc" and tok[0] == "switching_atom":
            tok[1] = str(int(tok[1]) - off)
        out.append(" ".join(tok))
    open(fep1, "w").write("\n".join(out) + "\n")
    want, keep0 = q.as_struct()
    py, keep1 = build_system(t, load_fep(fep1, t), use_LRF=bool(q.use_LRF)).as_struct()
    lib = host_lib()
    h = C.c_void_p()
    assert lib.qhost_open(top.encode(), fep1.encode(), int(q.use_LRF), -1, C.byref(h)) == 0, lib.qhost_last_error().decode()
    try:
        cs = lib.qhost_system(h).contents
        for st in (py, cs):
            assert st.qswitch == want.qswitch
            a, b = _struct_arrays(st), _struct_arrays(want)
            for k in a:
                assert np.array_equal(a[k], b[k]), k
    finally:
        lib.qhost_close(h)
