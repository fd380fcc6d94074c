"""ctypes binding of the C ABI (include/qnb.h) with the reference's procedure names.

``Qnb.make_pair_lists`` and ``Qnb.pot_energy_nonbonds`` take the arguments of the
Fortran procedures they replace (nonbondene.f90:749, potene.f90:320) and are what
the parity tests call, so that a test reads like the reference's call sites.
The shared library is the product; there is no Python or CPU fallback -- if
``libqnb.so`` is missing or no GPU is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .system import QSystem, qnb_system

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libqnb.so")

QNB_FLAG_MD = 1
QNB_FLAG_QQ = 2
QNB_FLAG_NO_ENERGY = 4
QNB_FLAG_D_IS_ZERO = 8
QNB_FLAG_SOLVENT_RESTRAINTS = 16
MAX_SHELLS = 8


class SolventRestraints(C.Structure):
    """qnb_solvent_restraints (include/qnb.h): restrain_solvent / watpol parameters as the host prepares them."""
    _fields_ = [("xwcent", C.c_double * 3), ("rwat", C.c_double), ("fk_wsphere", C.c_double), ("shift", C.c_double),
                ("Dwmz", C.c_double), ("awmz", C.c_double), ("fkwpol", C.c_double), ("wpol_restr", C.c_int32),
                ("nwpolr_shell", C.c_int32), ("rout", C.c_double * MAX_SHELLS), ("dr", C.c_double * MAX_SHELLS),
                ("cstb", C.c_double * MAX_SHELLS)]


def wat_shells(xwcent, rwat, fk_wsphere=60.0, Dwmz=None, awmz=None, Tfree=300.0, wpol_restr=True, fkwpol=20.0,
               wpolr_layer=3.0, drout=0.5, crgQtot=0.0, rho_wat=0.0335, mu_w=None, dielectric=80.0):
    """The host-side preparation of the solvent restraint parameters: defaults of md.f90 (fk_wsphere 60, fkwpol 20,
    Dwmz = 0.26 exp(-0.19 (rwat-15)) + 0.74, awmz = 0.2/(1+exp(0.4 (rwat-25))) + 0.3 (simprep.f90:4659-4671),
    wpolr_layer 3, drout 0.5) and wat_shells (simprep.f90:4745-4886:
    shells of width drout, 2 drout, ... from the surface inwards; cstb = crgQtot (1-1/eps) / (rho mu_w 4 pi rshell^2))."""
    import math
    p = SolventRestraints()
    for c in range(3):
        p.xwcent[c] = float(xwcent[c])
    p.rwat, p.fk_wsphere, p.fkwpol = float(rwat), float(fk_wsphere), float(fkwpol)
    p.awmz = float(awmz) if awmz is not None else 0.2 / (1.0 + math.exp(0.4 * (rwat - 25.0))) + 0.3
    p.Dwmz = float(Dwmz) if Dwmz is not None else 0.26 * math.exp(-0.19 * (rwat - 15.0)) + 0.74
    boltz = 0.001986
    p.shift = math.sqrt(boltz * Tfree / fk_wsphere) if fk_wsphere != 0.0 else 0.0
    p.wpol_restr = 1 if wpol_restr else 0
    drs = wpolr_layer / drout
    nsh = int(-0.5 + math.sqrt(2.0 * drs + 0.25))
    assert nsh <= MAX_SHELLS
    p.nwpolr_shell = nsh if wpol_restr else 0
    if mu_w is None:
        mu_w = 0.834 * 0.9572 * math.cos(math.radians(104.52) / 2.0)   # -chg_solv(1) * bnd0 * cos(ang0/2), TIP3P
    rout = rwat
    eps = 1.0 - 1.0 / dielectric
    for i in range(nsh):
        dr = drout * (i + 1)
        ri = rout - dr
        rshell = (0.5 * (rout ** 3 + ri ** 3)) ** (1.0 / 3.0)
        p.rout[i], p.dr[i] = rout, dr
        p.cstb[i] = crgQtot * eps / (rho_wat * mu_w * 4.0 * math.pi * rshell ** 2)
        rout -= dr
    return p
LIST_PP, LIST_PW, LIST_WW, LIST_QP, LIST_QW, LIST_QQ, LIST_QQP = range(7)
LRF_STRIDE = 43
IPC_BLOB = 128
E_COUNT = 7
EQ_STRIDE = 6
E_NAMES = ("pp.el", "pp.vdw", "pw.el", "pw.vdw", "ww.el", "ww.vdw", "LRF")
EQ_NAMES = ("qq.el", "qq.vdw", "qp.el", "qp.vdw", "qw.el", "qw.vdw")

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)
_PL = C.POINTER(C.c_int64)
_PF = C.POINTER(C.c_float)


class QnbError(RuntimeError):
    """A non-zero return of the C ABI; the Fortran wrapper would call die()."""


def bind(lib, prefix: str):
    """Declare the argument types shared by libqnb (prefix 'qnb') and the oracle (prefix 'qo')."""
    f = getattr(lib, f"{prefix}_last_error")
    f.restype = C.c_char_p
    f.argtypes = []
    return lib


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libqnb.so and declare every entry point of include/qnb.h."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise QnbError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA library is the product; there is no CPU fallback)")
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.qnb_last_error.restype = C.c_char_p
    lib.qnb_last_error.argtypes = []
    lib.qnb_device_count.restype = C.c_int
    lib.qnb_device_count.argtypes = []
    lib.qnb_init.restype = C.c_int
    lib.qnb_init.argtypes = [C.POINTER(qnb_system), C.c_int, C.POINTER(H)]
    lib.qnb_update_box.restype = C.c_int
    lib.qnb_update_box.argtypes = [H, _PD, _PD]
    for fn in (lib.qnb_save_lists, lib.qnb_restore_lists):
        fn.restype = C.c_int
        fn.argtypes = [H]
    lib.qnb_set_solvent_restraints.restype = C.c_int
    lib.qnb_set_solvent_restraints.argtypes = [H, C.POINTER(SolventRestraints)]
    lib.qnb_set_theta_corr.restype = C.c_int
    lib.qnb_set_theta_corr.argtypes = [H, _PD]
    lib.qnb_last_restraints.restype = C.c_int
    lib.qnb_last_restraints.argtypes = [H, _PD, _PD, _PI]
    lib.qnb_set_constraints.restype = C.c_int
    lib.qnb_set_constraints.argtypes = [H, C.c_int, _PI, _PI, _PD, _PD]
    lib.qnb_shake.restype = C.c_int
    lib.qnb_shake.argtypes = [H, _PD, _PD, _PL]
    lib.qnb_qcp_beads.restype = C.c_int
    lib.qnb_qcp_beads.argtypes = [H, _PD, C.c_int, _PI, C.c_int, _PD, _PD, _PD]
    lib.qnb_build_lists.restype = C.c_int
    lib.qnb_build_lists.argtypes = [H, _PD] + [C.c_double] * 7 + [_PL]
    lib.qnb_nonbond.restype = C.c_int
    # per-step call: raw addresses (array.ctypes costs microseconds per argument)
    lib.qnb_nonbond.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qnb_build_lists_batch.restype = C.c_int
    lib.qnb_build_lists_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 7 + [C.c_void_p]
    lib.qnb_nonbond_batch.restype = C.c_int
    lib.qnb_nonbond_batch.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
    lib.qnb_register_host_buffers.restype = C.c_int
    lib.qnb_register_host_buffers.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.qnb_release_host_buffers.restype = C.c_int
    lib.qnb_release_host_buffers.argtypes = [H]
    lib.qnb_list_count.restype = C.c_int
    lib.qnb_list_count.argtypes = [H, C.c_int, C.c_int, _PL]
    lib.qnb_export_list.restype = C.c_int
    lib.qnb_export_list.argtypes = [H, C.c_int, C.c_int, _PI, _PD, C.c_int64]
    lib.qnb_export_lrf.restype = C.c_int
    lib.qnb_export_lrf.argtypes = [H, _PD]
    lib.qnb_comm_unique_id.restype = C.c_int
    lib.qnb_comm_unique_id.argtypes = [C.c_void_p]
    lib.qnb_comm_init.restype = C.c_int
    lib.qnb_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.qnb_comm_ipc_export.restype = C.c_int
    lib.qnb_comm_ipc_export.argtypes = [H, C.c_void_p]
    lib.qnb_comm_status.restype = C.c_int
    lib.qnb_comm_status.argtypes = [H]
    lib.qnb_comm_ipc_attach.restype = C.c_int
    lib.qnb_comm_ipc_attach.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.qnb_bench_nonbond.restype = C.c_int
    lib.qnb_bench_nonbond.argtypes = [H, _PD, C.c_int, C.c_int, C.c_int, _PF]
    lib.qnb_bench_build_lists.restype = C.c_int
    lib.qnb_bench_build_lists.argtypes = [H, C.c_int, _PF]
    lib.qnb_bench_last_batch_timing.restype = C.c_int
    lib.qnb_bench_last_batch_timing.argtypes = [_PD]
    lib.qnb_bench_allreduce_phases.restype = C.c_int
    lib.qnb_bench_allreduce_phases.argtypes = [H, _PD]
    lib.qnb_bench_allreduce.restype = C.c_int
    lib.qnb_bench_allreduce.argtypes = [H, C.c_int, _PF]
    lib.qnb_bench_last_build_timing.restype = C.c_int
    lib.qnb_bench_last_build_timing.argtypes = [H, _PF]
    lib.qnb_bench_md.restype = C.c_int
    lib.qnb_bench_md.argtypes = [H, _PD, C.c_int, C.c_int, C.c_int, _PF]
    lib.qnb_bench_peak.restype = C.c_int
    lib.qnb_bench_peak.argtypes = [C.c_int, C.c_int, _PF]
    lib.qnb_bench_kernels.restype = C.c_int
    lib.qnb_bench_kernels.argtypes = [H, _PD, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, _PF, C.c_int]
    lib.qnb_launch_count.restype = C.c_int64
    lib.qnb_launch_count.argtypes = [H]
    lib.qnb_last_timing.restype = C.c_int
    lib.qnb_last_timing.argtypes = [H, _PD]
    lib.qnb_last_copy_bytes.restype = C.c_int
    lib.qnb_last_copy_bytes.argtypes = [H, _PL, _PL]
    lib.qnb_finalize.restype = C.c_int
    lib.qnb_finalize.argtypes = [H]
    if path == LIB_PATH:
        _lib = lib
    return lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_PD)


def _addr(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]


def _f64(a) -> np.ndarray:
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


class Qnb:
    """One handle = one GPU = one (shard of a) system."""

    def __init__(self, qsys: QSystem, device: int = 0, lib=None):
        self.lib = lib or load_library()
        self.sys = qsys
        if self.lib.qnb_device_count() <= 0:
            raise QnbError("no usable CUDA device: " + self.lib.qnb_last_error().decode())
        st, keep = qsys.as_struct()
        h = C.c_void_p()
        rc = self.lib.qnb_init(C.byref(st), device, C.byref(h))
        del keep
        if rc != 0:
            raise QnbError(self.lib.qnb_last_error().decode())
        self.h = h
        if qsys.use_PBC:
            self.update_box(qsys.boxlength)

    # -- error plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise QnbError(self.lib.qnb_last_error().decode())

    def register_host_buffers(self, x=None, d=None):
        """Page-lock the caller's persistent x / d arrays (qnb_register_host_buffers): later steps given exactly these
        arrays skip the staging copy of x and have the gradient added into d by the device.  The caller keeps the arrays
        alive until close() / release_host_buffers()."""
        for a in (x, d):
            assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.size == 3 * self.sys.natom)
        self._registered = (x, d)     # keep them alive as long as the handle
        self._check(self.lib.qnb_register_host_buffers(self.h, None if x is None else x.ctypes.data, None if d is None else d.ctypes.data))

    def release_host_buffers(self):
        self._check(self.lib.qnb_release_host_buffers(self.h))
        self._registered = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.qnb_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference-named operations
    def update_box(self, boxlength, inv_boxl=None):
        b = np.ascontiguousarray(boxlength, dtype=np.float64)
        ib = np.ascontiguousarray(1.0 / b if inv_boxl is None else inv_boxl, dtype=np.float64)
        self._check(self.lib.qnb_update_box(self.h, _dp(b), _dp(ib)))

    def save_lists(self):
        """MC_volume keeping the lists of the current box: old_nbww = nbww ... old_lrf = lrf (md.f90:2022-2058)."""
        self._check(self.lib.qnb_save_lists(self.h))

    def restore_lists(self):
        """MC_volume, rejected move: lists, LRF moments and box as saved (md.f90:2214-2256)."""
        self._check(self.lib.qnb_restore_lists(self.h))

    def make_pair_lists(self, x, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF=None, counts=True):
        """make_pair_lists(Rq,Rcq2,RcLRF2,Rcpp2,Rcpw2,Rcww2), nonbondene.f90:749.

        counts=True also returns nbpp_pair..nbqw_pair as the reference would log them under -DDUMP
        (md.f90:1693-1695); the exact nbpp count needs a host-side expansion, so production loops pass False."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert x.size == 3 * self.sys.natom
        if RcLRF is None:
            RcLRF = float(np.sqrt(RcLRF2)) if RcLRF2 >= 0 else -1.0
        out = np.zeros(8, np.int64) if counts else None
        self._check(self.lib.qnb_build_lists(self.h, _dp(x), Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF,
                                             out.ctypes.data_as(_PL) if counts else None))
        return out

    def set_solvent_restraints(self, params: SolventRestraints, theta_corr=None):
        """Hand restrain_solvent / watpol to the device step (pot_energy, potene.f90:161-167)."""
        self._check(self.lib.qnb_set_solvent_restraints(self.h, C.byref(params)))
        self._nshell = int(params.nwpolr_shell)
        if theta_corr is not None:
            self.set_theta_corr(theta_corr)

    def set_theta_corr(self, theta_corr):
        tc = np.ascontiguousarray(theta_corr, dtype=np.float64)
        assert tc.size >= getattr(self, "_nshell", 0)
        self._check(self.lib.qnb_set_theta_corr(self.h, _dp(tc)))

    def last_restraints(self):
        """(E[solvent_radial, water_pol], per-shell sum of theta, per-shell n_insh) of the last step with restraints."""
        n = max(getattr(self, "_nshell", 0), 1)
        E = np.zeros(2)
        ts = np.zeros(n)
        ns = np.zeros(n, np.int32)
        self._check(self.lib.qnb_last_restraints(self.h, _dp(E), _dp(ts), ns.ctypes.data_as(_PI)))
        k = getattr(self, "_nshell", 0)
        return E, ts[:k], ns[:k]

    def pot_energy_nonbonds(self, x, lambdas, md=True, qq=True, d=None, energies=True, restraints=False, d_is_zero=False):
        """pot_energy_nonbonds(E,EQ,md) (+ nonbond_qq/nonbond_qqp when qq): returns (d, E[7], EQ[nstates][6]).
        energies=False (extension, QNB_FLAG_NO_ENERGY): the pp/pw/ww energies of this step are not needed."""
        x = _f64(x)
        lam = _f64(lambdas)
        assert x.size == 3 * self.sys.natom and lam.size == self.sys.nstates
        if d is None:
            d = np.zeros((self.sys.natom, 3))
        else:
            assert d.dtype == np.float64 and d.flags.c_contiguous and d.size == 3 * self.sys.natom
        E = np.empty(E_COUNT)
        EQ = np.empty((self.sys.nstates, EQ_STRIDE))
        flags = ((QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0) | (0 if energies else QNB_FLAG_NO_ENERGY)
                 | (QNB_FLAG_SOLVENT_RESTRAINTS if restraints else 0) | (QNB_FLAG_D_IS_ZERO if d_is_zero else 0))
        if self.lib.qnb_nonbond(self.h, _addr(x), _addr(lam), flags, _addr(d), _addr(E), _addr(EQ)):
            self._check(1)
        return d.reshape(-1, 3), E, EQ

    def bind_step(self, x, lambdas, d, md=True, qq=True, d_is_zero=False):
        """The per-step call with everything bound once, for hosts whose x, lambda and d arrays live as long as the run (what a
        compiled host does by holding the pointers): returns (step, E, EQ) where step() runs qnb_nonbond on the CURRENT
        contents of x / lambdas, adds the gradient to d (or writes it: d_is_zero) and refreshes E[7], EQ[nstates][6]."""
        for a in (x, d):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == 3 * self.sys.natom
        assert lambdas.dtype == np.float64 and lambdas.size == self.sys.nstates
        E = np.empty(E_COUNT)
        EQ = np.empty((self.sys.nstates, EQ_STRIDE))
        flags = (QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0) | (QNB_FLAG_D_IS_ZERO if d_is_zero else 0)
        fn, h, args = self.lib.qnb_nonbond, self.h, (_addr(x), _addr(lambdas), flags, _addr(d), _addr(E), _addr(EQ))
        keep = (x, lambdas, d, E, EQ)

        def step(_keep=keep):
            if fn(h, *args):
                self._check(1)

        return step, E, EQ

    def qcp_beads(self, x_save, atoms, coord, lambdas):
        """qcp_run's bead loop (qcp.f90:319-372): for every bead, x(atoms) = x_save(atoms) + coord[bead] and
        pot_energy(..., .false.); returns EQ[nbeads][nstates][6] (nonbonded Q terms).  atoms are 1-based."""
        x = np.ascontiguousarray(x_save, dtype=np.float64).reshape(-1)
        at = np.ascontiguousarray(atoms, dtype=np.int32).reshape(-1)
        cd = np.ascontiguousarray(coord, dtype=np.float64)
        assert cd.ndim == 3 and cd.shape[1] == at.size and cd.shape[2] == 3
        lam = np.ascontiguousarray(lambdas, dtype=np.float64).reshape(-1)
        assert lam.size == self.sys.nstates
        out = np.zeros((cd.shape[0], self.sys.nstates, EQ_STRIDE))
        self._check(self.lib.qnb_qcp_beads(self.h, _dp(x), at.size, at.ctypes.data_as(_PI), cd.shape[0], _dp(cd), _dp(lam), _dp(out)))
        return out

    def list_count(self, which: int, state: int = 1) -> int:
        n = C.c_int64()
        self._check(self.lib.qnb_list_count(self.h, which, state, C.byref(n)))
        return n.value

    def export_list(self, which: int, state: int = 1, params: bool = True):
        n = self.list_count(which, state)
        ij = np.zeros((max(n, 1), 2), np.int32)
        p = np.zeros((max(n, 1), 4)) if params else None
        self._check(self.lib.qnb_export_list(self.h, which, state, ij.ctypes.data_as(_PI),
                                             _dp(p) if params else None, n))
        return ij[:n], (p[:n] if params else None)

    def export_lrf(self):
        out = np.zeros((self.sys.ncgp, LRF_STRIDE))
        self._check(self.lib.qnb_export_lrf(self.h, _dp(out)))
        return out

    # -- SHAKE of solvent-sized molecules (N2)
    def set_constraints(self, constraints, istart_mol, winv):
        """init_constraints' result (simprep.f90:2167-2345): rows (i, j, dist2) with 1-based atoms in topology order, grouped
        into molecules by istart_mol (1-based first atom of every molecule) like const_mol(:)."""
        cons = list(constraints)
        starts = np.asarray(istart_mol, np.int64)
        mol = np.searchsorted(starts, np.array([c[0] for c in cons], np.int64), side="right") - 1 if cons else np.zeros(0, np.int64)
        if len(mol) and (np.diff(mol) < 0).any():
            raise QnbError("constraints must be listed molecule by molecule")
        first = [0]
        for k in range(1, len(cons) + 1):
            if k == len(cons) or mol[k] != mol[k - 1]:
                first.append(k)
        first = np.ascontiguousarray(first, np.int32)
        ij = np.ascontiguousarray([[c[0], c[1]] for c in cons], np.int32).reshape(-1, 2)
        d2 = np.ascontiguousarray([c[2] for c in cons], np.float64)
        w = np.ascontiguousarray(winv, np.float64)
        if ij.size == 0:
            ij, d2 = np.zeros((1, 2), np.int32), np.zeros(1)
        self._check(self.lib.qnb_set_constraints(self.h, len(first) - 1, first.ctypes.data_as(_PI), ij.ctypes.data_as(_PI),
                                                 _dp(d2), _dp(w)))

    def shake(self, x, xx=None):
        """shake(xx, x), bondene.f90:1069: returns (constrained copy of x, sweeps summed over molecules).  xx=None: the
        coordinates resident on the device from the last pot_energy_nonbonds / make_pair_lists are the reference."""
        x = np.array(x, dtype=np.float64).reshape(-1)
        n = C.c_int64()
        xxp = None if xx is None else _dp(np.ascontiguousarray(xx, dtype=np.float64).reshape(-1))
        self._check(self.lib.qnb_shake(self.h, xxp, _dp(x), C.byref(n)))
        return x.reshape(-1, 3), n.value

    # -- multi-GPU
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.qnb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        self._check(self.lib.qnb_comm_init(self.h, rank, nranks, C.c_char_p(uid)))

    def ipc_export(self) -> bytes:
        """Descriptor of this rank's device buffers for the peer-memory all-reduce (qnb_comm_ipc_export)."""
        buf = C.create_string_buffer(IPC_BLOB)
        self._check(self.lib.qnb_comm_ipc_export(self.h, buf))
        return buf.raw

    def ipc_attach(self, rank: int, nranks: int, blobs):
        """Map the buffers of all ranks (their ipc_export() blobs, by rank): sums go over peer memory from now on."""
        raw = b"".join(blobs)
        assert len(raw) == nranks * IPC_BLOB
        self._check(self.lib.qnb_comm_ipc_attach(self.h, rank, nranks, C.c_char_p(raw)))

    def comm_status(self):
        self._check(self.lib.qnb_comm_status(self.h))

    def comm_connect(self, rank: int, nranks: int, dist, mode: str = "auto"):
        """What the host does once per run: exchange the descriptors (or the NCCL id) over its own process group
        (MPI_Allgather / MPI_Bcast in the Fortran host, torch.distributed here).  mode: 'p2p' (peer memory), 'nccl',
        'auto' (peer memory when the GPUs can reach each other, else NCCL).  Returns the mode in use."""
        if mode in ("p2p", "auto"):
            blobs = [None] * nranks
            dist.all_gather_object(blobs, self.ipc_export())
            try:
                self.ipc_attach(rank, nranks, blobs)
                ok = 1
            except QnbError:
                if mode == "p2p":
                    raise
                ok = 0
            oks = [None] * nranks
            dist.all_gather_object(oks, ok)
            if all(oks):
                return "p2p"
            if mode == "p2p":
                raise QnbError("peer-memory attach failed on another rank")
        uid = [self.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        self.comm_init(rank, nranks, uid[0])
        return "nccl"

    # -- measurement
    def bench_nonbond(self, lambdas, steps: int, md=True, qq=True, flush_l2=False, energies=True, restraints=False) -> float:
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        ms = C.c_float()
        flags = ((QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0) | (0 if energies else QNB_FLAG_NO_ENERGY)
                 | (QNB_FLAG_SOLVENT_RESTRAINTS if restraints else 0))
        self._check(self.lib.qnb_bench_nonbond(self.h, _dp(lam), flags, steps, int(flush_l2), C.byref(ms)))
        return ms.value

    def bench_md(self, lambdas, steps: int, nbcycle: int, md=True, qq=True) -> float:
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        ms = C.c_float()
        flags = (QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0)
        self._check(self.lib.qnb_bench_md(self.h, _dp(lam), flags, steps, nbcycle, C.byref(ms)))
        return ms.value

    def bench_build_lists(self, reps: int) -> float:
        ms = C.c_float()
        self._check(self.lib.qnb_bench_build_lists(self.h, reps, C.byref(ms)))
        return ms.value

    def bench_allreduce(self, reps: int) -> float:
        ms = C.c_float()
        self._check(self.lib.qnb_bench_allreduce(self.h, reps, C.byref(ms)))
        return ms.value

    def allreduce_phases(self) -> dict:
        t = np.zeros(3)
        self._check(self.lib.qnb_bench_allreduce_phases(self.h, _dp(t)))
        return dict(zip(("wait_arrive_us", "sum_deliver_us", "wait_done_us"), t.tolist()))

    def last_build_timing(self) -> dict:
        """ms per build of the parts timed by the last bench_build_lists: LRF accumulation, row scan passes."""
        t = (C.c_float * 3)()
        self._check(self.lib.qnb_bench_last_build_timing(self.h, t))
        return dict(zip(("lrf_ms", "rows_count_ms", "rows_fill_ms"), [float(v) for v in t]))

    def bench_kernels(self, lambdas, reps: int, md=True, qq=True, flush_l2=False, restraints=False) -> dict:
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        flags = (QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0) | (QNB_FLAG_SOLVENT_RESTRAINTS if restraints else 0)
        n = self.lib.qnb_bench_kernels(self.h, _dp(lam), flags, reps, int(flush_l2), names, 4096, ms, 64)
        if n < 0:
            raise QnbError(self.lib.qnb_last_error().decode())
        nm = names.value.decode().split("\n")[:n]
        return {nm[i]: ms[i] for i in range(n)}

    def launch_count(self) -> int:
        return int(self.lib.qnb_launch_count(self.h))

    def last_timing(self) -> dict:
        t = np.zeros(4)
        self._check(self.lib.qnb_last_timing(self.h, _dp(t)))
        return dict(zip(("stage_in_s", "issue_s", "device_wait_s", "add_out_s"), t.tolist()))

    def last_copy_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.qnb_last_copy_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


class QnbBatch:
    """Independent systems of one process (the lambda windows of a FEP farm, EVB frames) advanced together on one GPU
    through qnb_build_lists_batch / qnb_nonbond_batch: same arguments and results as the single-system calls, one list
    entry per system.  The pointer tables are built once; the arrays they point to are reused every step."""

    def __init__(self, handles):
        assert handles and len({id(g) for g in handles}) == len(handles)
        self.g = list(handles)
        self.lib = self.g[0].lib
        n = self.n = len(self.g)
        self._h = (C.c_void_p * n)(*[g.h for g in self.g])
        self.x = [np.zeros(3 * g.sys.natom) for g in self.g]
        self.lam = [np.zeros(g.sys.nstates) for g in self.g]
        self.d = [np.zeros((g.sys.natom, 3)) for g in self.g]
        self.E = [np.zeros(E_COUNT) for _ in self.g]
        self.EQ = [np.zeros((g.sys.nstates, EQ_STRIDE)) for g in self.g]
        for g, xk, dk in zip(self.g, self.x, self.d):
            g.release_host_buffers()
            g.register_host_buffers(xk.reshape(-1), dk.reshape(-1))     # the batch owns these arrays for its lifetime
        tab = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        self._x, self._lam, self._d, self._E, self._EQ = tab(self.x), tab(self.lam), tab(self.d), tab(self.E), tab(self.EQ)

    def _check(self, rc):
        if rc:
            raise QnbError(self.lib.qnb_last_error().decode())

    def make_pair_lists(self, xs, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF=None, counts=False):
        for k, xk in enumerate(xs):
            self.x[k][:] = np.asarray(xk, dtype=np.float64).reshape(-1)
        if RcLRF is None:
            RcLRF = float(np.sqrt(RcLRF2)) if RcLRF2 >= 0 else -1.0
        out = np.zeros((self.n, 8), np.int64) if counts else None
        self._check(self.lib.qnb_build_lists_batch(self.n, self._h, self._x, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF,
                                                   out.ctypes.data if counts else None))
        return out

    def pot_energy_nonbonds(self, xs=None, lambdas=None, md=True, qq=True, zero_d=True):
        """One step of every system; xs / lambdas (lists) are copied into the batch's own arrays when given.  Returns the
        batch's (d, E, EQ) lists: d[k] is ADDED to unless zero_d."""
        if xs is not None:
            for k, xk in enumerate(xs):
                self.x[k][:] = np.asarray(xk, dtype=np.float64).reshape(-1)
        if lambdas is not None:
            for k, lk in enumerate(lambdas):
                self.lam[k][:] = lk
        # zero_d: d(:) = zero of pot_energy (potene.f90:109) -- the library is told (QNB_FLAG_D_IS_ZERO) and writes the
        # gradient into the registered arrays, so the clearing itself is not needed
        flags = (QNB_FLAG_MD if md else 0) | (QNB_FLAG_QQ if qq else 0) | (QNB_FLAG_D_IS_ZERO if zero_d else 0)
        if self.lib.qnb_nonbond_batch(self.n, self._h, self._x, self._lam, flags, self._d, self._E, self._EQ):
            self._check(1)
        return self.d, self.E, self.EQ

    def last_timing(self) -> dict:
        t = np.zeros(3)
        self._check(self.lib.qnb_bench_last_batch_timing(_dp(t)))
        return dict(zip(("issue_s", "device_wait_s", "add_out_s"), t.tolist()))


def bench_peak(which: int, device: int = 0) -> float:
    """Measured FMA throughput (TFLOP/s) of the FP32 (0) or FP64 (1) pipe."""
    lib = load_library()
    tf = C.c_float()
    if lib.qnb_bench_peak(device, which, C.byref(tf)) != 0:
        raise QnbError(lib.qnb_last_error().decode())
    return tf.value
