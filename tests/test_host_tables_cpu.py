"""The PRODUCT's host-side tables (q6_b200/csrc/qnb_tables.hpp: HostTables::build, what qnb_init runs on the CPU before
anything is uploaded) against the oracle, without a GPU.

HostTables is the library's own statement of precompute_interactions (simprep.f90:2860-3589) and nbqqlist
(nonbondene.f90:3231-3282): per-type LJ data + a sparse special-pair table instead of pp_map/pp_precomp, dense Q tables,
the static nbqq/nbqqp lists.  For every entry of every pair list the oracle builds, the parameters the product would use
(vdWA, vdWB, elec, score) must be the oracle's -- and the input checks of qnb_init must refuse what the kernels do not
support.  tests/cpu_shims/host_tables_shim.cpp (test infrastructure, compiled here with g++) exposes the header.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from common import golden_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SRC = os.path.join(ROOT, "tests", "cpu_shims", "host_tables_shim.cpp")
SHIM_LIB = os.path.join(ROOT, "tests", "cpu_shims", "libhost_tables_shim.so")
HDR = os.path.join(ROOT, "q6_b200", "csrc", "qnb_tables.hpp")

_PD, _PI = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def shim():
    from q6_b200.system import qnb_system
    newest = max(os.path.getmtime(p) for p in (SHIM_SRC, HDR, os.path.join(ROOT, "include", "qnb.h")))
    if not os.path.exists(SHIM_LIB) or os.path.getmtime(SHIM_LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SHIM_LIB, SHIM_SRC])
    lib = C.CDLL(SHIM_LIB)
    lib.ht_build.restype = C.c_void_p
    lib.ht_build.argtypes = [C.POINTER(qnb_system), C.c_char_p, C.c_int]
    lib.ht_free.argtypes = [C.c_void_p]
    lib.ht_pp_many.argtypes = [C.c_void_p, C.c_long, _PI, _PD, _PI, _PI]
    lib.ht_qp_many.argtypes = [C.c_void_p, C.c_long, _PI, C.c_int, _PD, _PI]
    lib.ht_pp.argtypes = [C.c_void_p, C.c_int, C.c_int, _PD, _PI]
    lib.ht_pw.argtypes = [C.c_void_p, C.c_int, C.c_int, _PD]
    lib.ht_ww.argtypes = [C.c_void_p, C.c_int, C.c_int, _PD]
    lib.ht_qw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _PD]
    lib.ht_static_count.restype = C.c_long
    lib.ht_static_count.argtypes = [C.c_void_p, C.c_int]
    lib.ht_static.argtypes = [C.c_void_p, C.c_int, C.c_long, _PI, _PD]
    lib.ht_atom_flags.argtypes = [C.c_void_p, C.c_int]
    return lib


def _build(lib, q, **override):
    st, keep = q.as_struct()
    for k, v in override.items():
        setattr(st, k, v)
    err = C.create_string_buffer(512)
    h = lib.ht_build(C.byref(st), err, 512)
    return h, err.value.decode(), keep


def _cases():
    out = [(n,) + tuple(golden_system(n)[:3]) for n in ("c1_sph", "c1_pbc", "c4_evb")]
    names = ("sph_fep2", "sph_evb2", "sph_small_rcq", "sph_nowater", "box_solute_q", "box_solute_rowimage", "sph_anyatom")
    return out + [c for c in common.small_systems() if c[0] in names] + common.variant_systems()


def _same(a, b):
    return np.allclose(a, b, rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_product_host_tables_match_oracle_lists(shim, case):
    from oracle.pyoracle import Oracle
    name, q, cuts, lam = case
    h, err, keep = _build(shim, q)
    assert h, err
    o = Oracle(q)
    try:
        o.make_pair_lists(q.xtop, **cuts)
        ns, sa = q.nat_solute, q.solv_atom
        # nbpp: every listed pair is known to the product as not excluded, %set, with the same parameters
        ij, p = o.export_list(0, 1)
        if len(ij):
            ijc = np.ascontiguousarray(ij, np.int32)
            out, st, ok = np.zeros((len(ij), 4)), np.zeros(len(ij), np.int32), np.zeros(len(ij), np.int32)
            shim.ht_pp_many(h, len(ij), ijc.ctypes.data_as(_PI), out.ctypes.data_as(_PD), st.ctypes.data_as(_PI),
                            ok.ctypes.data_as(_PI))
            assert ok.all() and st.all()
            assert _same(out, p)
        # nbpw / nbww: parameters depend on (solute atom, water site) / (site, site) only
        ij, p = o.export_list(1, 1)
        if len(ij):
            site = (ij[:, 1] - ns - 1) % sa
            key = ij[:, 0].astype(np.int64) * sa + site
            u, first = np.unique(key, return_index=True)
            buf = np.zeros(4)
            for k, f in zip(u, first):
                shim.ht_pw(h, int(k // sa) - 1, int(k % sa), buf.ctypes.data_as(_PD))
                assert _same(buf, p[f]), (k, buf, p[f])
            assert _same(p, p[first][np.searchsorted(u, key)])      # and the oracle is consistent over the list
        ij, p = o.export_list(2, 1)
        if len(ij):
            a, b = (ij[:, 0] - ns - 1) % sa, (ij[:, 1] - ns - 1) % sa
            buf = np.zeros(4)
            for aa in range(sa):
                for bb in range(sa):
                    m = (a == aa) & (b == bb)
                    shim.ht_ww(h, aa, bb, buf.ctypes.data_as(_PD))
                    assert m.any() and _same(p[m], buf[None, :])
        for s in range(1, q.nstates + 1):
            # nbqp: (Q-atom, atom) per state
            ij, p = o.export_list(3, s)
            if len(ij):
                ijc = np.ascontiguousarray(ij, np.int32)
                out, st = np.zeros((len(ij), 4)), np.zeros(len(ij), np.int32)
                shim.ht_qp_many(h, len(ij), ijc.ctypes.data_as(_PI), s - 1, out.ctypes.data_as(_PD), st.ctypes.data_as(_PI))
                assert _same(out, p)
                # listed partners are non-Q atoms further than three bonds from every Q-atom
                fl = np.array([shim.ht_atom_flags(h, int(a) - 1) for a in np.unique(ij[:, 1])])
                assert not (fl & 3).any()
            # nbqw: (Q-atom, water atom)
            ij, p = o.export_list(4, s)
            if len(ij):
                site = (ij[:, 1] - ns - 1) % sa
                key = ij[:, 0].astype(np.int64) * sa + site
                u, first = np.unique(key, return_index=True)
                buf = np.zeros(4)
                for k, f in zip(u, first):
                    shim.ht_qw(h, int(k // sa) - 1, s - 1, int(k % sa), buf.ctypes.data_as(_PD))
                    assert _same(buf, p[f])
            # static nbqq / nbqqp: the product's lists ARE the oracle's, entry for entry
            for which, lst in ((0, 5), (1, 6)):
                ij, p = o.export_list(lst, s)
                mine = []
                ids, buf = np.zeros(6, np.int32), np.zeros(4)
                for k in range(shim.ht_static_count(h, which)):
                    shim.ht_static(h, which, k, ids.ctypes.data_as(_PI), buf.ctypes.data_as(_PD))
                    if ids[4] == s - 1:
                        mine.append((int(ids[2]), int(ids[3]) if which == 0 else int(ids[1]) + 1) + tuple(buf))
                mine.sort()
                ref = sorted((int(a), int(b)) + tuple(pp) for (a, b), pp in zip(ij, p))
                assert len(mine) == len(ref)
                for m, r in zip(mine, ref):
                    assert m[:2] == r[:2] and _same(np.array(m[2:]), np.array(r[2:])), (m, r)
    finally:
        o.close()
        shim.ht_free(h)


def test_product_classifies_exclusions_and_14_pairs(shim):
    """pp_code against the topology's own tables: listex / listexlong pairs are excluded, list14 / list14long pairs get
    LJ code 3 and el14_scale, everything else the ljcod type pair (c4_evb has all four kinds)."""
    q, cuts, lam, z = golden_system("c4_evb")
    h, err, keep = _build(shim, q)
    assert h, err
    try:
        ns, R = q.nat_solute, q.max_nbr_range
        nonq = np.asarray(q.iqatom)[:ns] == 0
        buf, st = np.zeros(4), C.c_int()
        listex, list14 = np.asarray(q.listex).reshape(ns, R), np.asarray(q.list14).reshape(ns, R)
        crg = np.asarray(q.crg)
        n_ex = n_14 = 0
        for i, k in zip(*np.nonzero(listex)):
            j = i + k + 1
            if j < ns and nonq[i] and nonq[j]:
                assert shim.ht_pp(h, int(i), int(j), buf.ctypes.data_as(_PD), C.byref(st)) == 0
                n_ex += 1
        for i, k in zip(*np.nonzero(list14 & ~listex.astype(bool))):
            j = i + k + 1
            if j < ns and nonq[i] and nonq[j]:
                assert shim.ht_pp(h, int(i), int(j), buf.ctypes.data_as(_PD), C.byref(st)) == 1
                assert np.isclose(buf[2], crg[i] * crg[j] * q.el14_scale, rtol=1e-14)
                n_14 += 1
        for a, b in np.asarray(q.listexlong).reshape(-1, 2):
            if abs(a - b) > R and nonq[a - 1] and nonq[b - 1]:
                assert shim.ht_pp(h, int(a) - 1, int(b) - 1, buf.ctypes.data_as(_PD), C.byref(st)) == 0
        assert n_ex > 1000 and n_14 > 1000
        # a plain pair: full charge product
        i, j = [int(v) for v in np.nonzero(nonq)[0][[0, -1]]]
        assert shim.ht_pp(h, i, j, buf.ctypes.data_as(_PD), C.byref(st)) == 1
        assert buf[2] == crg[i] * crg[j]
    finally:
        shim.ht_free(h)


@pytest.mark.parametrize("change,message", [
    (dict(solv_atom=4), "3-site solvent"),
    (dict(ntors_gt_solute=1), "internal torsions"),
    (dict(nstates=9), "nstates"),
    (dict(qswitch=0, use_PBC=1), "qswitch"),
])
def test_product_refuses_unsupported_systems(shim, change, message):
    """What qnb_init refuses (DESIGN.md 'out of scope'): n-site solvents, solvents with torsions, > 8 states, a
    periodic FEP system without switching atom -- loudly, with a message, never a silent fallback."""
    q, cuts, lam, z = golden_system("c1_sph")
    h, err, keep = _build(shim, q, **change)
    assert not h
    assert message in err, err
