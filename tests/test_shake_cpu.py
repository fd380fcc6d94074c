"""The algorithm of the device SHAKE on the CPU: q6_b200/csrc/qnb_shake.cuh's shake_molecule() is __host__ __device__, so
the source the kernel k_shake runs per thread is compiled here with g++ (tests/cpu_shims/shake_shim.cpp) and compared,
bit for bit, with the restatement of shake(xx, x) (bondene.f90:1069-1150) that the reference's step-0 goldens pin.
What this cannot cover is the thread mapping of the kernel and the copies in qnb_shake (tests/test_zx_shake_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from common import golden_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu_shims", "shake_shim.cpp")
HDR = os.path.join(ROOT, "q6_b200", "csrc", "qnb_shake.cuh")
LIB = os.path.join(ROOT, "tests", "cpu_shims", "libshake_shim.so")
_PD, _PI = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.shake_shim.argtypes = [C.c_int, _PI, _PI, _PD, _PD, _PD, _PD, C.POINTER(C.c_longlong)]
    return lib


def _run(lib, cons, nmol_first, winv, xx, x):
    ij0 = np.ascontiguousarray([[c[0] - 1, c[1] - 1] for c in cons], np.int32)
    d2 = np.ascontiguousarray([c[2] for c in cons], np.float64)
    first = np.ascontiguousarray(nmol_first, np.int32)
    xx = np.ascontiguousarray(xx, np.float64).reshape(-1)
    x = np.array(x, np.float64).reshape(-1)
    w = np.ascontiguousarray(winv, np.float64)
    n = C.c_longlong()
    failed = lib.shake_shim(len(first) - 1, first.ctypes.data_as(_PI), ij0.ctypes.data_as(_PI), d2.ctypes.data_as(_PD),
                            w.ctypes.data_as(_PD), xx.ctypes.data_as(_PD), x.ctypes.data_as(_PD), C.byref(n))
    return x.reshape(-1, 3), n.value, failed


def _waters(q, b_oh, b_hh):
    cons = []
    for k in range(q.nwat):
        o = q.nat_solute + 3 * k + 1
        cons += [(o, o + 1, b_oh ** 2), (o, o + 2, b_oh ** 2), (o + 1, o + 2, b_hh ** 2)]
    starts = ([1] if q.nat_solute else []) + [q.nat_solute + 3 * k + 1 for k in range(q.nwat)]
    mass = np.asarray(q.iaclib).reshape(-1, 7)[np.asarray(q.iac) - 1, 0]
    return cons, starts, 1.0 / mass, [3 * k for k in range(q.nwat + 1)]


@pytest.mark.parametrize("name", ["c1_sph", "c1_pbc"])
def test_kernel_source_reproduces_reference_step0_coordinates(shim, name):
    """The shipped systems: topology coordinates -> the fixture's post-SHAKE coordinates (the ones at which the
    reference's step-0 energies come out to the printed digit), sweep count included."""
    q, cuts, lam, z = golden_system(name)
    cons, starts, winv, first = _waters(q, 0.957, 1.5183)       # lig_w.top bond codes 13 / 14
    x, n, failed = _run(shim, cons, first, winv, q.xtop, q.xtop)
    assert not failed
    assert np.array_equal(x, z["x_step0"])
    assert n == int(z["shake_iterations"].sum())


def test_kernel_source_matches_restatement_on_displaced_coordinates(shim):
    """The leap-frog case (xx != x) on a synthetic system, against oracle.pyoracle.shake."""
    from oracle import pyoracle
    from q6_b200 import synth
    name, q, cuts, lam = [c for c in common.small_systems() if c[0] == "sph_fep2"][0]
    hh = 2.0 * synth.R_OH * float(np.sin(synth.ANG_HOH / 2))
    cons, starts, winv, first = _waters(q, synth.R_OH, hh)
    x0, n0, f0 = _run(shim, cons, first, winv, q.xtop, q.xtop)
    want0, nits0 = pyoracle.shake(cons, starts, winv, q.xtop, q.xtop)
    assert not f0 and n0 == sum(nits0) and np.array_equal(x0, want0)
    moved = x0 + np.random.default_rng(8).normal(0, 0.01, x0.shape)
    x1, n1, f1 = _run(shim, cons, first, winv, x0, moved)
    want1, nits1 = pyoracle.shake(cons, starts, winv, x0, moved)
    assert not f1 and n1 == sum(nits1) and np.array_equal(x1, want1)


def test_kernel_source_reports_failure(shim):
    """A constraint that cannot be met along the reference vector (xx perpendicular to x's bond): CONST_MAX_ITER sweeps,
    failure flag -- the caller turns it into die('shake failure')."""
    xx = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    x = np.array([[0.0, 0.0, 0.0], [1e-9, 3.0, 0.0]])
    out, n, failed = _run(shim, [(1, 2, 1.0)], [0, 1], np.ones(2), xx, x)
    assert failed == 1 and n == 1000
