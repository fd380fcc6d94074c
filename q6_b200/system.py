"""Host-side preparation of the static tables the nonbonded engine consumes.

In the reference these steps stay in Fortran on the host and run before the
nonbonded path is first entered; they are mirrored here (Python, because no
Fortran toolchain exists in this image) so that the shipped topology / FEP
fixtures can drive the C ABI exactly as ``qdyn.f90:113-153`` would:

* ``topology``        simprep.f90:4521-4571  (ljcod table, sqrt(eps) for the arithmetic rule)
* ``get_fep``         simprep.f90:997-1182   (iqatom, default qcrg, removal of redefined bonds,
                                               special exclusions folded into listex/listexlong)
* ``prep_sim``        simprep.f90:3592-4000  (charge scaling by sqrt(coulomb_constant))
* ``make_qconn``      nonbondene.f90:3087-3187
* ``distribute_nonbonds`` nonbondene.f90:80-505 (i-range assignment; equal shares)

The result, :class:`QSystem`, is a field-for-field image of ``qnb_system``
(include/qnb.h) with numpy arrays in the Fortran host's layout.
"""
from __future__ import annotations

import ctypes as C
from collections import deque
from dataclasses import dataclass, field

import numpy as np

from .fep import Fep
from .topo import MAX_NBR_RANGE, Topology

QNB_ABI_VERSION = 1
I32 = np.int32
F64 = np.float64


class qnb_system(C.Structure):
    """ctypes image of ``struct qnb_system`` (include/qnb.h)."""
    _I = C.c_int32
    _PI = C.POINTER(C.c_int32)
    _PD = C.POINTER(C.c_double)
    _fields_ = (
        [(n, C.c_int32) for n in (
            "abi_version", "natom", "nat_solute", "nwat", "solv_atom", "ncgp", "ncgp_solute", "nqat", "nstates",
            "qswitch", "natyps", "num_atyp", "max_nbr_range", "nexlong", "n14long", "nqlib", "nqexpnb", "nel_scale",
            "iuse_switch_atom", "use_PBC", "use_LRF", "ivdw_rule", "solvent_type", "qvdw_flag",
            "qq_use_library_charges", "ntors_gt_solute")]
        + [("el14_scale", C.c_double), ("xpcent", C.c_double * 3), ("rexcl_o", C.c_double)]
        + [("cgp", _PI), ("cgpatom", _PI), ("excl", _PI), ("iqatom", _PI), ("iqseq", _PI), ("iac", _PI),
           ("crg", _PD), ("iaclib", _PD), ("ljcod", _PI), ("listex", _PI), ("list14", _PI), ("listexlong", _PI),
           ("list14long", _PI),
           ("qcrg", _PD), ("qiac", _PI), ("qavdw", _PD), ("qbvdw", _PD), ("sc_lookup", _PD), ("iqexpnb", _PI),
           ("jqexpnb", _PI), ("el_scale_iq", _PI), ("el_scale_jq", _PI), ("el_scale", _PD), ("qconn", _PI)]
        + [(n, C.c_int32) for n in (
            "pp_start", "pp_end", "pw_start", "pw_end", "qp_start", "qp_end", "ww_start", "ww_end", "qw_start",
            "qw_end", "natom_start", "natom_end", "is_master")]
    )


def _i32(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=I32).reshape(-1))
    return a if a.size or n is None else a


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=F64).reshape(-1))


@dataclass
class QSystem:
    natom: int = 0
    nat_solute: int = 0
    nwat: int = 0
    solv_atom: int = 3
    ncgp: int = 0
    ncgp_solute: int = 0
    nqat: int = 0
    nstates: int = 1
    qswitch: int = 0
    natyps: int = 0
    num_atyp: int = 0
    max_nbr_range: int = MAX_NBR_RANGE
    nqlib: int = 0
    iuse_switch_atom: int = 1
    use_PBC: int = 0
    use_LRF: int = 1
    ivdw_rule: int = 1
    solvent_type: int = 0
    qvdw_flag: int = 0
    qq_use_library_charges: int = 0
    ntors_gt_solute: int = 0
    el14_scale: float = 1.0
    xpcent: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rexcl_o: float = 0.0
    boxlength: np.ndarray = field(default_factory=lambda: np.zeros(3))
    # arrays, flattened in Fortran (column-major) order
    cgp: np.ndarray = None        # [ncgp][3] iswitch, first, last
    cgpatom: np.ndarray = None
    excl: np.ndarray = None
    iqatom: np.ndarray = None
    iqseq: np.ndarray = None
    iac: np.ndarray = None
    crg: np.ndarray = None        # scaled
    iaclib: np.ndarray = None     # [natyps][7]
    ljcod: np.ndarray = None      # [num_atyp][num_atyp] (symmetric)
    listex: np.ndarray = None     # [nat_solute][max_nbr_range]  == Fortran (k,i) column-major
    list14: np.ndarray = None
    listexlong: np.ndarray = None  # [n][2]
    list14long: np.ndarray = None
    qcrg: np.ndarray = None       # [nqat][nstates] scaled (python order; flattened to (iq,state) col-major on export)
    qiac: np.ndarray = None       # [nqat][nstates]
    qavdw: np.ndarray = None      # [nqlib][3]
    qbvdw: np.ndarray = None
    sc_lookup: np.ndarray = None  # [nqat][natyps+nqat][nstates]
    iqexpnb: np.ndarray = None
    jqexpnb: np.ndarray = None
    el_scale_iq: np.ndarray = None
    el_scale_jq: np.ndarray = None
    el_scale: np.ndarray = None   # [n][nstates]
    qconn: np.ndarray = None      # [nqat][nat_solute][nstates]  (== Fortran (state,atom,iq) col-major)
    # shard
    pp: tuple = (1, 0)
    pw: tuple = (1, 0)
    qp: tuple = (1, 0)
    ww: tuple = (1, 0)
    qw: tuple = (1, 0)
    natom_range: tuple = (1, 0)
    is_master: int = 1
    # topology coordinates (convenience for tests/bench)
    xtop: np.ndarray = None

    _SCALARS = ("natom", "nat_solute", "nwat", "solv_atom", "ncgp", "ncgp_solute", "nqat", "nstates", "qswitch",
                "natyps", "num_atyp", "max_nbr_range", "nqlib", "iuse_switch_atom", "use_PBC", "use_LRF", "ivdw_rule",
                "solvent_type", "qvdw_flag", "qq_use_library_charges", "ntors_gt_solute", "el14_scale", "rexcl_o")
    _ARRAYS = ("xpcent", "boxlength", "cgp", "cgpatom", "excl", "iqatom", "iqseq", "iac", "crg", "iaclib", "ljcod",
               "listex", "list14", "listexlong", "list14long", "qcrg", "qiac", "qavdw", "qbvdw", "sc_lookup",
               "iqexpnb", "jqexpnb", "el_scale_iq", "el_scale_jq", "el_scale", "qconn", "xtop")

    def save(self, path: str) -> None:
        """Fixture file: every table of the system (compressed npz)."""
        d = {k: np.asarray(getattr(self, k)) for k in self._SCALARS}
        for k in self._ARRAYS:
            v = getattr(self, k)
            d[k] = np.zeros(0) if v is None else np.asarray(v)
        # the exclusion bit tables are sparse: store packed
        d["listex"] = np.packbits(np.asarray(self.listex, np.uint8).reshape(-1))
        d["list14"] = np.packbits(np.asarray(self.list14, np.uint8).reshape(-1))
        d["qconn"] = np.asarray(self.qconn, np.int8)
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "QSystem":
        z = np.load(path)
        q = QSystem()
        for k in QSystem._SCALARS:
            v = z[k]
            setattr(q, k, float(v) if k in ("el14_scale", "rexcl_o") else int(v))
        for k in QSystem._ARRAYS:
            setattr(q, k, z[k])
        n = q.nat_solute * q.max_nbr_range
        q.listex = np.unpackbits(z["listex"])[:n].astype(np.int32).reshape(q.nat_solute, q.max_nbr_range)
        q.list14 = np.unpackbits(z["list14"])[:n].astype(np.int32).reshape(q.nat_solute, q.max_nbr_range)
        q.qconn = z["qconn"].astype(np.int32)
        return q.full_shard()

    def full_shard(self) -> "QSystem":
        """calculation_assignment for numnodes == 1 (nonbondene.f90:126-152)."""
        self.pp = self.pw = self.qp = (1, self.ncgp_solute)
        self.ww = self.qw = (1, self.nwat)
        self.natom_range = (1, self.natom)
        self.is_master = 1
        return self

    def as_struct(self):
        """Returns (qnb_system, keepalive) -- keepalive must outlive the struct's use."""
        keep = []

        def pi(a):
            a = np.ascontiguousarray(np.asarray(a if a is not None else [], dtype=I32).reshape(-1))
            if a.size == 0:
                a = np.zeros(1, I32)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32))

        def pd(a):
            a = np.ascontiguousarray(np.asarray(a if a is not None else [], dtype=F64).reshape(-1))
            if a.size == 0:
                a = np.zeros(1, F64)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double))

        s = qnb_system()
        s.abi_version = QNB_ABI_VERSION
        for n in ("natom", "nat_solute", "nwat", "solv_atom", "ncgp", "ncgp_solute", "nqat", "nstates", "qswitch",
                  "natyps", "num_atyp", "max_nbr_range", "nqlib", "iuse_switch_atom", "use_PBC", "use_LRF",
                  "ivdw_rule", "solvent_type", "qvdw_flag", "qq_use_library_charges", "ntors_gt_solute"):
            setattr(s, n, int(getattr(self, n)))
        s.nexlong = 0 if self.listexlong is None else int(np.asarray(self.listexlong).reshape(-1, 2).shape[0])
        s.n14long = 0 if self.list14long is None else int(np.asarray(self.list14long).reshape(-1, 2).shape[0])
        s.nqexpnb = 0 if self.iqexpnb is None else int(np.asarray(self.iqexpnb).size)
        s.nel_scale = 0 if self.el_scale_iq is None else int(np.asarray(self.el_scale_iq).size)
        s.el14_scale = float(self.el14_scale)
        s.xpcent = (C.c_double * 3)(*[float(v) for v in self.xpcent])
        s.rexcl_o = float(self.rexcl_o)
        s.cgp = pi(self.cgp)
        s.cgpatom = pi(self.cgpatom)
        s.excl = pi(self.excl)
        s.iqatom = pi(self.iqatom)
        s.iqseq = pi(self.iqseq)
        s.iac = pi(self.iac)
        s.crg = pd(self.crg)
        s.iaclib = pd(self.iaclib)
        s.ljcod = pi(self.ljcod)
        s.listex = pi(self.listex)
        s.list14 = pi(self.list14)
        s.listexlong = pi(self.listexlong)
        s.list14long = pi(self.list14long)
        nq, ns = self.nqat, self.nstates
        # python [iq][state] -> Fortran (iq,state) column-major == transpose then C-flatten
        s.qcrg = pd(np.asarray(self.qcrg if self.qcrg is not None else np.zeros((nq, ns))).reshape(nq, ns).T)
        s.qiac = pi(np.asarray(self.qiac if self.qiac is not None else np.zeros((nq, ns), I32)).reshape(nq, ns).T)
        nl = self.nqlib
        s.qavdw = pd(np.asarray(self.qavdw if self.qavdw is not None else np.zeros((nl, 3))).reshape(nl, 3).T)
        s.qbvdw = pd(np.asarray(self.qbvdw if self.qbvdw is not None else np.zeros((nl, 3))).reshape(nl, 3).T)
        nk = self.natyps + nq
        sc = np.asarray(self.sc_lookup if self.sc_lookup is not None else np.zeros((nq, nk, ns))).reshape(nq, nk, ns)
        s.sc_lookup = pd(sc.transpose(2, 1, 0))  # (iq,k,state) col-major
        s.iqexpnb = pi(self.iqexpnb)
        s.jqexpnb = pi(self.jqexpnb)
        s.el_scale_iq = pi(self.el_scale_iq)
        s.el_scale_jq = pi(self.el_scale_jq)
        ne = s.nel_scale
        s.el_scale = pd(np.asarray(self.el_scale if self.el_scale is not None else np.zeros((ne, ns))).reshape(ne, ns).T)
        s.qconn = pi(self.qconn)  # stored [iq][atom][state] == (state,atom,iq) col-major
        s.pp_start, s.pp_end = self.pp
        s.pw_start, s.pw_end = self.pw
        s.qp_start, s.qp_end = self.qp
        s.ww_start, s.ww_end = self.ww
        s.qw_start, s.qw_end = self.qw
        s.natom_start, s.natom_end = self.natom_range
        s.is_master = int(self.is_master)
        return s, keep


def make_qconn(nstates, nat_solute, nqat, iqseq, iqatom, bnd_solute, qbnd_ij, qbnd_cod, exspec_ij, exspec_flag):
    """``make_qconn``/``find_bonded`` (nonbondene.f90:3087-3187).

    qconn(state, atom, iq) = 1 for the Q-atom itself, 1 + number of bonds to it up to 4,
    9 further away; walks topology bonds with cod != 0 and the state's Q-bonds
    (``qbnd%cod(state) > 0``).  The reference's depth-first search keeps the minimum
    level, i.e. the breadth-first distance computed here.  Special exclusions write 0.
    Returned as [iq][atom][state] (C order) == Fortran (state,atom,iq).
    """
    qconn = np.full((nqat, nat_solute, nstates), 9, dtype=I32)
    adj0 = [[] for _ in range(nat_solute + 1)]
    for i, j, cod in np.asarray(bnd_solute).reshape(-1, 3):
        if cod == 0:
            continue
        if i <= nat_solute and j <= nat_solute:
            adj0[i].append(j)
            adj0[j].append(i)
    for s in range(nstates):
        adj = [list(a) for a in adj0]
        for b, (i, j) in enumerate(np.asarray(qbnd_ij).reshape(-1, 2)):
            if qbnd_cod[b][s] > 0:
                adj[i].append(j)
                adj[j].append(i)
        for iq in range(nqat):
            src = int(iqseq[iq])
            level = {src: 1}
            dq = deque([src])
            while dq:
                cur = dq.popleft()
                if level[cur] >= 4:
                    continue
                for nb in adj[cur]:
                    if nb not in level:
                        level[nb] = level[cur] + 1
                        dq.append(nb)
            for a, lv in level.items():
                qconn[iq, a - 1, s] = lv
    for k, (i, j) in enumerate(np.asarray(exspec_ij).reshape(-1, 2)):
        for a, b in ((i, j), (j, i)):
            iq = iqatom[a - 1]
            if iq > 0:
                for s in range(nstates):
                    if exspec_flag[k][s]:
                        qconn[iq - 1, b - 1, s] = 0
    return qconn


def build_system(topo: Topology, fep: Fep | None, use_LRF: bool = True) -> QSystem:
    """topology + get_fep + prep_sim + make_qconn for one node (numnodes == 1 assignment)."""
    q = QSystem()
    q.natom, q.nat_solute, q.solv_atom, q.nwat = topo.nat_pro, topo.nat_solute, topo.solv_atom, topo.nwat
    q.ncgp, q.ncgp_solute, q.iuse_switch_atom = topo.ncgp, topo.ncgp_solute, topo.iuse_switch_atom
    q.natyps, q.ivdw_rule, q.el14_scale = topo.natyps, topo.ivdw_rule, topo.el14_scale
    q.solvent_type, q.use_PBC, q.use_LRF = topo.solvent_type, int(topo.use_PBC), int(use_LRF)
    q.xpcent, q.rexcl_o, q.boxlength = topo.xpcent.copy(), topo.rexcl_o, topo.boxlength.copy()
    q.ntors_gt_solute = int(topo.ntors > topo.ntors_solute)
    q.cgp, q.cgpatom, q.iac = topo.cgp.copy(), topo.cgpatom.copy(), topo.iac.copy()
    q.excl = topo.excl.copy()
    q.xtop = topo.xtop.copy()
    # topology(): ljcod and sqrt(eps) (simprep.f90:4553-4571)
    q.num_atyp = int(topo.iac.max()) if topo.iac.size else 0
    lj = np.ones((q.num_atyp, q.num_atyp), I32)
    for i, j in topo.lj2:
        lj[i - 1, j - 1] = 2
        lj[j - 1, i - 1] = 2
    q.ljcod = lj
    iaclib = topo.iaclib.copy()
    if topo.ivdw_rule == 2:
        iaclib[:, 4:7] = np.sqrt(np.abs(iaclib[:, 4:7]))
    q.iaclib = iaclib
    listex, list14 = topo.listex.copy(), topo.list14.copy()
    listexlong = [tuple(r) for r in topo.listexlong]
    q.list14long = topo.list14long.copy()
    crg = topo.crg.copy()
    bnd = topo.bnd[:topo.nbonds_solute].copy()

    q.iqatom = np.zeros(q.natom, I32)
    if fep is not None and fep.nqat > 0:
        q.nqat, q.nstates, q.qswitch = fep.nqat, fep.nstates, fep.qswitch
        q.iqseq = fep.iqseq.copy()
        for i, a in enumerate(fep.iqseq):  # get_fep L1010-1019
            if 0 < a <= q.nat_solute:
                q.iqatom[a - 1] = i + 1
        q.qvdw_flag, q.qq_use_library_charges = int(fep.qvdw_flag), int(fep.qq_use_library_charges)
        q.nqlib, q.qiac = fep.nqlib, fep.qiac.copy()
        q.qavdw, q.qbvdw = fep.qavdw.copy(), fep.qbvdw.copy()
        if fep.qvdw_flag and topo.ivdw_rule == 2:  # get_fep L1041-1046
            q.qbvdw[:, 0] = np.sqrt(q.qbvdw[:, 0])
            q.qbvdw[:, 2] = np.sqrt(q.qbvdw[:, 2])
        q.sc_lookup = fep.sc_lookup.copy()
        q.iqexpnb, q.jqexpnb = fep.iqexpnb.copy(), fep.jqexpnb.copy()
        q.el_scale_iq, q.el_scale_jq, q.el_scale = fep.el_scale_iq.copy(), fep.el_scale_jq.copy(), fep.el_scale.copy()
        # redefined bonds are switched off in the topology (get_fep L1049-1065)
        for b in range(bnd.shape[0]):
            for (qi, qj) in fep.qbnd_ij:
                if (bnd[b, 0] == qi and bnd[b, 1] == qj) or (bnd[b, 0] == qj and bnd[b, 1] == qi):
                    bnd[b, 2] = 0
        # special exclusions involving a non-Q atom go into the exclusion lists (get_fep L1143-1179)
        for k, (i, j) in enumerate(fep.exspec_ij):
            if i < 1 or i > q.natom or j < 1 or j > q.natom:
                raise ValueError("invalid special exclusion data")
            if q.iqatom[i - 1] == 0 or q.iqatom[j - 1] == 0:
                fl = fep.exspec_flag[k]
                if fl.any():
                    if not fl.all():
                        raise ValueError("Non-Q-atom special excl. pair must be on in all or no states")
                    if abs(j - i) <= MAX_NBR_RANGE:
                        if i < j:
                            listex[i - 1, j - i - 1] = 1
                        else:
                            listex[j - 1, i - j - 1] = 1
                    else:
                        listexlong.append((i, j))
        q.qconn = make_qconn(q.nstates, q.nat_solute, q.nqat, q.iqseq, q.iqatom, bnd, fep.qbnd_ij, fep.qbnd_cod,
                             fep.exspec_ij, fep.exspec_flag)
        qcrg = fep.qcrg.copy()
    else:
        q.nqat, q.nstates = 0, (fep.nstates if fep is not None else 1)
        q.iqseq = np.zeros(0, I32)
        q.qconn = np.zeros(0, I32)
        q.sc_lookup = np.zeros((0, q.natyps, q.nstates))
        qcrg = np.zeros((0, q.nstates))
    q.listex, q.list14 = listex, list14
    q.listexlong = np.array(listexlong, dtype=I32).reshape(-1, 2)
    # prep_sim: charges scaled by sqrt(coulomb_constant) (simprep.f90:3714-3721)
    sq = np.sqrt(topo.coulomb_constant)
    q.crg = crg * sq
    q.qcrg = qcrg * sq
    return q.full_shard()


def distribute_nonbonds(per_i_counts: np.ndarray, nranks: int) -> list:
    """Contiguous i-ranges balanced by per-i pair counts (rule of distribute_nonbonds,
    nonbondene.f90:171-296, with equal shares: no master discount).  Returns 1-based
    inclusive (start, end) per rank; the union is 1..n, ranges may be empty."""
    n = len(per_i_counts)
    total = int(np.sum(per_i_counts))
    out = []
    i = 0
    acc = 0
    for r in range(nranks):
        start = i + 1
        if r == nranks - 1:
            i = n
        else:
            target = (total * (r + 1)) // nranks
            while i < n and acc < target:
                acc += int(per_i_counts[i])
                i += 1
        out.append((start, i))
    return out


def shard_system(q: QSystem, rank: int, nranks: int, pp_counts=None, ww_counts=None) -> QSystem:
    """A copy of q whose calculation_assignment is rank's share."""
    import copy
    s = copy.copy(q)
    ppc = np.ones(q.ncgp_solute) if pp_counts is None else pp_counts
    wwc = np.ones(q.nwat) if ww_counts is None else ww_counts
    pr = distribute_nonbonds(ppc, nranks)[rank]
    wr = distribute_nonbonds(wwc, nranks)[rank]
    s.pp = s.pw = s.qp = pr
    s.ww = s.qw = wr
    s.natom_range = distribute_nonbonds(np.ones(q.natom), nranks)[rank]
    s.is_master = int(rank == 0)
    return s
