"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/libqoracle.so (see qoracle.h).

Same method names as q6_b200.engine.Qnb so parity tests call both sides alike.
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from q6_b200.system import QSystem, qnb_system

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqoracle.so")
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)
_PL = C.POINTER(C.c_int64)
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qoracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "qoracle.h")),
            os.path.getmtime(os.path.join(_HERE, "..", "include", "qnb.h"))):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libqoracle.so"])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.qo_last_error.restype = C.c_char_p
    lib.qo_create.restype = H
    lib.qo_create.argtypes = [C.POINTER(qnb_system)]
    lib.qo_destroy.argtypes = [H]
    lib.qo_update_box.argtypes = [H, _PD, _PD]
    lib.qo_make_pair_lists.restype = C.c_int
    lib.qo_make_pair_lists.argtypes = [H, _PD] + [C.c_double] * 7 + [_PL]
    lib.qo_nonbond.restype = C.c_int
    lib.qo_nonbond.argtypes = [H, _PD, _PD, C.c_int, _PD, _PD, _PD]
    lib.qo_list_count.restype = C.c_int
    lib.qo_list_count.argtypes = [H, C.c_int, C.c_int, _PL]
    lib.qo_export_list.restype = C.c_int
    lib.qo_export_list.argtypes = [H, C.c_int, C.c_int, _PI, _PD, C.c_int64]
    lib.qo_export_lrf.restype = C.c_int
    lib.qo_export_lrf.argtypes = [H, _PD]
    lib.qo_make_qconn.restype = None
    lib.qo_make_qconn.argtypes = [C.c_int, C.c_int, C.c_int, _PI, _PI, C.c_int, _PI, C.c_int, _PI, _PI, C.c_int,
                                  _PI, _PI, _PI]
    lib.qo_solvent_restraints.restype = C.c_int
    lib.qo_solvent_restraints.argtypes = [C.POINTER(qnb_system), C.c_void_p, _PD, _PD, C.c_int, _PD, _PD, _PD, _PI]
    lib.qo_time_decomposed.restype = C.c_double
    lib.qo_time_decomposed.argtypes = [C.POINTER(qnb_system), _PD, _PD, C.c_int, _PD, _PD, C.c_int, C.c_int,
                                       C.c_int, _PD, _PD, _PD, _PD]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(_PD)


class Oracle:
    def __init__(self, qsys: QSystem):
        self.lib = load()
        self.sys = qsys
        st, keep = qsys.as_struct()
        self.h = self.lib.qo_create(C.byref(st))
        if not self.h:
            raise RuntimeError(self.lib.qo_last_error().decode())
        if qsys.use_PBC:
            self.update_box(qsys.boxlength)

    def close(self):
        if getattr(self, "h", None):
            self.lib.qo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def update_box(self, boxlength, inv_boxl=None):
        b = np.ascontiguousarray(boxlength, dtype=np.float64)
        ib = np.ascontiguousarray(1.0 / b if inv_boxl is None else inv_boxl, dtype=np.float64)
        self.lib.qo_update_box(self.h, _dp(b), _dp(ib))

    def make_pair_lists(self, x, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF=None):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if RcLRF is None:
            RcLRF = float(np.sqrt(RcLRF2)) if RcLRF2 >= 0 else -1.0
        counts = np.zeros(8, np.int64)
        rc = self.lib.qo_make_pair_lists(self.h, _dp(x), Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF,
                                         counts.ctypes.data_as(_PL))
        if rc:
            raise RuntimeError(self.lib.qo_last_error().decode())
        return counts

    def pot_energy_nonbonds(self, x, lambdas, md=True, qq=True, d=None):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        lam = np.ascontiguousarray(lambdas, dtype=np.float64).reshape(-1)
        if d is None:
            d = np.zeros(3 * self.sys.natom)
        E = np.zeros(7)
        EQ = np.zeros(6 * self.sys.nstates)
        flags = (1 if md else 0) | (2 if qq else 0)
        self.lib.qo_nonbond(self.h, _dp(x), _dp(lam), flags, _dp(d.reshape(-1)), _dp(E), _dp(EQ))
        return d.reshape(-1, 3), E, EQ.reshape(self.sys.nstates, 6)

    def list_count(self, which, state=1):
        n = C.c_int64()
        self.lib.qo_list_count(self.h, which, state, C.byref(n))
        return n.value

    def export_list(self, which, state=1, params=True):
        n = self.list_count(which, state)
        ij = np.zeros((max(n, 1), 2), np.int32)
        p = np.zeros((max(n, 1), 4))
        self.lib.qo_export_list(self.h, which, state, ij.ctypes.data_as(_PI), _dp(p), n)
        return ij[:n], (p[:n] if params else None)

    def export_lrf(self):
        out = np.zeros((self.sys.ncgp, 43))
        self.lib.qo_export_lrf(self.h, _dp(out))
        return out


def solvent_restraints(qsys: QSystem, params, theta_corr, x, md=True):
    """restrain_solvent + watpol (nonbondene.f90:6466-6746): returns (d, E[2], shell_theta_sum, shell_n)."""
    lib = load()
    st, keep = qsys.as_struct()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    tc = np.ascontiguousarray(theta_corr, dtype=np.float64)
    d = np.zeros(3 * qsys.natom)
    E = np.zeros(2)
    n = max(int(params.nwpolr_shell), 1)
    ts, ns = np.zeros(n), np.zeros(n, np.int32)
    rc = lib.qo_solvent_restraints(C.byref(st), C.addressof(params), _dp(tc), _dp(x), int(md), _dp(d), _dp(E), _dp(ts),
                                   ns.ctypes.data_as(_PI))
    if rc:
        raise RuntimeError("qo_solvent_restraints failed")
    k = int(params.nwpolr_shell)
    return d.reshape(-1, 3), E, ts[:k], ns[:k]


CONST_TOL = 0.0001          # globals.f90:519
CONST_MAX_ITER = 1000       # globals.f90:520


def shake_constraints(bnd, bondlib, mass, nat_solute, shake_solvent=True, shake_solute=False, shake_hydrogens=True,
                      shake_heavy=False):
    """init_constraints (simprep.f90:2167-2345) restated for the bond part: the constrained bonds in topology order as
    rows (i, j, dist2) with 1-based atom numbers.  heavy = mass >= 4 (topo.f90:896-903); bonds with code 0 are
    skipped; defaults are those of md.f90:338-357 (solvent bonds to hydrogens: O-H, O-H and the H-H "bond" of TIP3P)."""
    heavy = np.asarray(mass) >= 4.0
    out = []
    for ia, ja, cod in np.asarray(bnd).reshape(-1, 3):
        if cod == 0:
            continue
        has_h = (not heavy[ia - 1]) or (not heavy[ja - 1])
        on = (shake_solute and ia <= nat_solute) or (shake_solvent and ia > nat_solute)
        if on and ((has_h and shake_hydrogens) or (not has_h and shake_heavy)):
            out.append((int(ia), int(ja), float(bondlib[cod - 1][1]) ** 2))
    return out


def shake(constraints, istart_mol, winv, xx, x):
    """SHAKE as ``shake(xx, x)`` of bondene.f90:1069-1150: x is corrected along the bond vectors of the reference
    coordinates xx until every constraint of a molecule is within CONST_TOL (relative, on the squared length).  Literal
    restatement incl. the order of operations: a constraint found in range is flagged ready and still receives that
    pass's correction.  Returns (x_new, iterations per molecule)."""
    x = np.array(x, dtype=np.float64)
    xx = np.asarray(xx, dtype=np.float64)
    bounds = list(istart_mol) + [len(x) + 1]
    by_mol = {}
    for c in constraints:
        m = int(np.searchsorted(bounds, c[0], side="right")) - 1
        by_mol.setdefault(m, []).append(c)
    nits_all = []
    for m in sorted(by_mol):
        cons = by_mol[m]
        ready = [False] * len(cons)
        nits = 0
        while True:
            for ic, (i, j, dist2) in enumerate(cons):
                if ready[ic]:
                    continue
                i0, j0 = i - 1, j - 1
                xij = x[i0] - x[j0]                          # q_dist5(x(j), x(i))%vec, math.f90:270
                diff = dist2 - (xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2])
                if abs(diff) < CONST_TOL * dist2:
                    ready[ic] = True
                xxij = xx[i0] - xx[j0]
                scp = xij[0] * xxij[0] + xij[1] * xxij[1] + xij[2] * xxij[2]
                corr = diff / (2.0 * scp * (winv[i0] + winv[j0]))
                x[i0] = x[i0] + xxij * corr * winv[i0]
                x[j0] = x[j0] + (-xxij) * corr * winv[j0]
            nits += 1
            if all(ready):
                break
            if nits >= CONST_MAX_ITER:
                raise RuntimeError("shake failure")          # die('shake failure'), bondene.f90:1143
        nits_all.append(nits)
    return x, nits_all


def initial_constraint_x(topo, x=None):
    """The coordinate part of initial_constraint (bondene.f90:1035-1036; called from qdyn.f90:133 when iseed > 0):
    xx = x; shake(xx, x).  The velocity part does not touch x.  These are the coordinates the reference's step-0
    energies (the first row of tests/basic_tests/*_benchmark.en) are evaluated at."""
    x = topo.xtop if x is None else x
    mass = topo.iaclib[np.asarray(topo.iac) - 1, 0]
    cons = shake_constraints(topo.bnd, topo.bondlib, mass, topo.nat_solute)
    return shake(cons, topo.istart_mol, 1.0 / mass, x, x)


def make_qconn(nstates, nat_solute, nqat, iqseq, iqatom, bnd_solute, qbnd_ij, qbnd_cod, exspec_ij, exspec_flag):
    """Literal find_bonded recursion (nonbondene.f90:3131); returns [iq][atom][state]."""
    lib = load()
    I = np.int32
    iqseq = np.ascontiguousarray(iqseq, I)
    iqatom = np.ascontiguousarray(iqatom, I)
    bnd = np.ascontiguousarray(bnd_solute, I).reshape(-1, 3)
    qb = np.ascontiguousarray(qbnd_ij, I).reshape(-1, 2)
    qc = np.ascontiguousarray(np.asarray(qbnd_cod, I).reshape(-1, nstates).T)  # (bond,state) col-major
    ex = np.ascontiguousarray(exspec_ij, I).reshape(-1, 2)
    ef = np.ascontiguousarray(np.asarray(exspec_flag, I).reshape(-1, nstates).T)
    out = np.zeros(nstates * nat_solute * nqat, I)
    pi = lambda a: (a if a.size else np.zeros(1, I)).ctypes.data_as(_PI)
    lib.qo_make_qconn(nstates, nat_solute, nqat, pi(iqseq), pi(iqatom), bnd.shape[0], pi(bnd), qb.shape[0], pi(qb),
                      pi(qc), ex.shape[0], pi(ex), pi(ef), out.ctypes.data_as(_PI))
    return out.reshape(nqat, nat_solute, nstates)


def time_decomposed(qsys: QSystem, x, lambdas, cut7, nthreads, steps, md=True, qq=True, build_each=False):
    """Qdyn6p-style decomposed CPU run; returns dict(seconds, list_seconds, d, E, EQ)."""
    lib = load()
    st, keep = qsys.as_struct()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    lam = np.ascontiguousarray(lambdas, dtype=np.float64)
    cut = np.ascontiguousarray(cut7, dtype=np.float64)
    box = None
    if qsys.use_PBC:
        box = np.concatenate([qsys.boxlength, 1.0 / qsys.boxlength]).astype(np.float64)
    d = np.zeros(3 * qsys.natom)
    E = np.zeros(7)
    EQ = np.zeros(6 * qsys.nstates)
    ls = C.c_double()
    flags = (1 if md else 0) | (2 if qq else 0)
    sec = lib.qo_time_decomposed(C.byref(st), _dp(x), _dp(lam), flags, _dp(cut), _dp(box) if box is not None else None,
                                 nthreads, steps, int(build_each), _dp(d), _dp(E), _dp(EQ),
                                 C.cast(C.byref(ls), _PD))
    if sec < 0:
        raise RuntimeError(lib.qo_last_error().decode())
    return dict(seconds=sec, list_seconds=ls.value, d=d.reshape(-1, 3), E=E, EQ=EQ.reshape(qsys.nstates, 6))
