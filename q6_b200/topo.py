"""Q topology file reader (host side, stays on the CPU in the reference too).

Mirrors ``topo_read`` (reference src/topo.f90:522-1114): the same record order,
the same list-directed / fixed-format reads, the same optional fields and
defaults.  Only what the nonbonded path and ``make_qconn`` consume is kept; the
bonded libraries are parsed (they must be, to stay in step with the records) and
dropped.  All index arrays keep the file's 1-based values.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MAX_NBR_RANGE = 25  # topo.f90:37
NLJTYP = 3  # topo.f90 (avdw/bvdw codes 1..3)


class TopologyError(RuntimeError):
    """topo_read's ``>>>>> ERROR: Could not read topology file.`` (topo.f90:1106)."""


def _to_float(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


class _Records:
    """Fortran sequential formatted input: every READ starts on a fresh record."""

    def __init__(self, text: str):
        self.lines = text.split("\n")
        self.pos = 0

    def line(self) -> str:
        if self.pos >= len(self.lines):
            raise TopologyError("unexpected end of topology file")
        s = self.lines[self.pos]
        self.pos += 1
        return s

    def skip(self) -> None:
        self.line()

    @staticmethod
    def leading(line: str, conv, nmax: int) -> list:
        """List-directed read of up to nmax leading values; stops at the first bad token."""
        out = []
        for tok in line.replace(",", " ").split():
            if len(out) == nmax:
                break
            try:
                out.append(conv(tok))
            except ValueError:
                break
        return out

    def values(self, n: int, conv) -> list:
        """One list-directed READ of n items (spans records, drops the rest of the last one)."""
        out: list = []
        while len(out) < n:
            for tok in self.line().replace(",", " ").split():
                if len(out) == n:
                    break
                out.append(conv(tok))
        return out

    def fixed_chars(self, n: int, width: int = 80) -> str:
        """Fixed-format READ of n one-character fields, `width` per record (80i1 / 80l1)."""
        out = []
        left = n
        while left > 0:
            rec = self.line()
            take = min(left, width)
            out.append(rec[:take].ljust(take))
            left -= take
        return "".join(out)


@dataclass
class Topology:
    title: str = ""
    version: float = 2.0
    nat_pro: int = 0
    nat_solute: int = 0
    solv_atom: int = 3
    nwat: int = 0
    xtop: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    iac: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    nbonds: int = 0
    nbonds_solute: int = 0
    bnd: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))  # i, j, cod
    bondlib: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))  # fk, bnd0 per bond code (topo.f90:691-700)
    nangles: int = 0
    nangles_solute: int = 0
    ntors: int = 0
    ntors_solute: int = 0
    nimps: int = 0
    nimps_solute: int = 0
    crg: np.ndarray = field(default_factory=lambda: np.zeros(0))
    ncgp: int = 0
    ncgp_solute: int = 0
    iuse_switch_atom: int = 1
    cgp: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))  # iswitch, first, last
    cgpatom: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    natyps: int = 0
    ivdw_rule: int = 1
    el14_scale: float = 1.0
    coulomb_constant: float = 332.0
    iaclib: np.ndarray = field(default_factory=lambda: np.zeros((0, 7)))  # mass, avdw(3), bvdw(3)
    lj2: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    list14: np.ndarray = field(default_factory=lambda: np.zeros((0, MAX_NBR_RANGE), np.int32))  # [atom][k]
    listex: np.ndarray = field(default_factory=lambda: np.zeros((0, MAX_NBR_RANGE), np.int32))
    list14long: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    listexlong: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    nres: int = 0
    nres_solute: int = 0
    res_start: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    res_name: list = field(default_factory=list)
    nmol: int = 0
    istart_mol: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tac: list = field(default_factory=list)
    solvent_type: int = 0
    use_PBC: bool = False
    boxlength: np.ndarray = field(default_factory=lambda: np.zeros(3))
    boxcentre: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rexcl_o: float = 0.0
    rwat: float = 0.0
    xpcent: np.ndarray = field(default_factory=lambda: np.zeros(3))
    xwcent: np.ndarray = field(default_factory=lambda: np.zeros(3))
    excl: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))


def topo_read(path: str) -> Topology:
    """``topo_read`` (topo.f90:522-1114)."""
    with open(path, "r", errors="replace") as fh:
        rec = _Records(fh.read())
    t = Topology()
    try:
        _read(rec, t)
    except (ValueError, IndexError) as e:  # the reference's err=1000 path
        raise TopologyError(f">>>>> ERROR: Could not read topology file. ({e})") from e
    return t


def _read(rec: _Records, t: Topology) -> None:
    # 1. header (L569-605)
    title = rec.line()[:80]
    if title.strip() == "Q topology file":
        while True:
            line = rec.line()
            parts = line.split(None, 1)
            key = parts[0].upper() if parts else ""
            rest = parts[1].strip() if len(parts) > 1 else ""
            if key == "TITLE":
                title = rest
            elif key == "VERSION":
                # strtod of the leading "a.b" part of the version string (L583-586)
                fields = rest.split()[0].split(".")
                t.version = float(".".join(fields[:2]))
            elif key == "END":
                break
    else:
        t.version = 2.0
    t.title = title

    # 2. nat_pro, nat_solute, solv_atom (L612-651)
    line = rec.line()
    v = rec.leading(line, int, 4)
    if not v:
        raise ValueError("nat_pro")
    t.nat_pro = v[0]
    if len(v) >= 4:
        t.nat_solute, t.solv_atom = v[1], v[2]
    elif len(v) == 3:
        t.nat_solute, t.solv_atom = v[1], v[2]
    elif len(v) == 2:
        t.nat_solute, t.solv_atom = v[1], 3
    else:
        t.nat_solute, t.solv_atom = t.nat_pro, 3
    if t.solv_atom == 0:
        t.nwat, t.solv_atom = 0, 1
    else:
        t.nwat = (t.nat_pro - t.nat_solute) // t.solv_atom

    # 3. coordinates (L661)
    if t.nat_pro > 0:
        t.xtop = np.array(rec.values(3 * t.nat_pro, _to_float), dtype=np.float64).reshape(t.nat_pro, 3)
    # 4. atom codes (L668-669)
    rec.skip()
    if t.nat_pro > 0:
        t.iac = np.array(rec.values(t.nat_pro, int), dtype=np.int32)

    # 5. bonds (L676-707)
    v = rec.leading(rec.line(), int, 2)
    t.nbonds = v[0]
    t.nbonds_solute = v[1] if len(v) > 1 else t.nbonds
    if t.nbonds > 0:
        t.bnd = np.array(rec.values(3 * t.nbonds, int), dtype=np.int32).reshape(t.nbonds, 3)
    nbndcod = rec.leading(rec.line(), int, 1)[0]
    t.bondlib = np.zeros((nbndcod, 2))
    for i in range(nbndcod):
        v = rec.leading(rec.line(), _to_float, 3)     # code number, fk, bnd0 [, SYBYL type]
        if len(v) >= 3:
            t.bondlib[i] = v[1:3]
    # 6. angles (L712-743)
    v = rec.leading(rec.line(), int, 2)
    t.nangles = v[0]
    t.nangles_solute = v[1] if len(v) > 1 else t.nangles
    if t.nangles > 0:
        rec.values(4 * t.nangles, int)
    nangcod = rec.leading(rec.line(), int, 1)[0]
    for _ in range(nangcod):
        rec.skip()
    # 7. torsions (L748-769)
    v = rec.leading(rec.line(), int, 2)
    t.ntors = v[0]
    t.ntors_solute = v[1] if len(v) > 1 else t.ntors
    if t.ntors > 0:
        rec.values(5 * t.ntors, int)
    ntorcod = rec.leading(rec.line(), int, 1)[0]
    for _ in range(ntorcod):
        rec.values(5, _to_float)
    # 8. impropers (L773-796)
    v = rec.leading(rec.line(), int, 2)
    t.nimps = v[0]
    t.nimps_solute = v[1] if len(v) > 1 else t.nimps
    if t.nimps > 0:
        rec.values(5 * t.nimps, int)
    nimpcod = rec.leading(rec.line(), int, 1)[0]
    for _ in range(nimpcod):
        rec.values(3, _to_float)

    # 9. charges (L801-804)
    v = rec.leading(rec.line(), int, 1)
    if v:
        t.nat_pro = v[0]
    if t.nat_pro > 0:
        t.crg = np.array(rec.values(t.nat_pro, _to_float), dtype=np.float64)

    # 10. charge groups (L811-832)
    v = rec.leading(rec.line(), int, 3)
    t.ncgp = v[0]
    t.ncgp_solute = v[1] if len(v) > 1 else t.ncgp
    t.iuse_switch_atom = v[2] if len(v) > 2 else 1
    cgp = np.zeros((t.ncgp, 3), np.int32)
    cgpatom = []
    k = 1
    for i in range(t.ncgp):
        n, isw = rec.values(2, int)
        cgp[i] = (isw, k, k + n - 1)
        cgpatom.extend(rec.values(n, int))
        k += n
    t.cgp = cgp
    t.cgpatom = np.array(cgpatom, dtype=np.int32)

    # 11. masses and vdW parameters (L847-891)
    t.natyps = rec.values(1, int)[0]
    t.ivdw_rule = rec.values(1, int)[0]
    v = rec.leading(rec.line(), _to_float, 2)
    t.el14_scale = v[0]
    t.coulomb_constant = v[1] if len(v) > 1 and v[1] > 0 else 332.0
    iaclib = np.zeros((t.natyps, 7))
    rec.skip()
    iaclib[:, 0] = rec.values(t.natyps, _to_float)
    for j in range(NLJTYP):
        rec.skip()
        iaclib[:, 1 + j] = rec.values(t.natyps, _to_float)
        rec.skip()
        iaclib[:, 4 + j] = rec.values(t.natyps, _to_float)
    t.iaclib = iaclib
    nlj2 = rec.values(1, int)[0]
    t.lj2 = np.array([rec.values(2, int) for _ in range(nlj2)], dtype=np.int32).reshape(nlj2, 2)

    # 12. 1-4 neighbour and exclusion lists (L912-963)
    rec.values(1, int)  # n14nbrs
    if t.nat_solute > 0:
        s = rec.fixed_chars(MAX_NBR_RANGE * t.nat_solute)
        t.list14 = (np.frombuffer(s.encode(), dtype=np.uint8) == ord("1")).astype(np.int32).reshape(
            t.nat_solute, MAX_NBR_RANGE)
    n14long = rec.values(1, int)[0]
    t.list14long = np.array([rec.values(2, int) for _ in range(n14long)], dtype=np.int32).reshape(n14long, 2)
    rec.values(1, int)  # nexnbrs
    if t.nat_solute > 0:
        s = rec.fixed_chars(MAX_NBR_RANGE * t.nat_solute)
        t.listex = (np.frombuffer(s.encode(), dtype=np.uint8) == ord("1")).astype(np.int32).reshape(
            t.nat_solute, MAX_NBR_RANGE)
    nexlong = rec.values(1, int)[0]
    t.listexlong = np.array([rec.values(2, int) for _ in range(nexlong)], dtype=np.int32).reshape(nexlong, 2)

    # residue / molecule bookkeeping (L966-995)
    line = rec.line()
    if line.strip() == "":
        line = rec.line()
    v = rec.leading(line, int, 2)
    t.nres = v[0]
    t.nres_solute = v[1] if len(v) > 1 else t.nres
    if t.nres > 0:
        t.res_start = np.array(rec.values(t.nres, int), dtype=np.int32)
    rec.skip()
    names = []
    for i in range((t.nres + 15) // 16 if t.nres > 0 else 0):
        line = rec.line()
        for j in range(min(16, t.nres - i * 16)):
            names.append(line[5 * j:5 * j + 4])
    t.res_name = names
    t.nmol = rec.values(1, int)[0]
    if t.nmol > 0:
        t.istart_mol = np.array(rec.values(t.nmol, int), dtype=np.int32)
    # atom type names (L997-1031)
    rec.skip()
    tac = []
    for i in range((t.natyps + 7) // 8 if t.natyps > 0 else 0):
        line = rec.line()
        for j in range(min(8, t.natyps - i * 8)):
            tac.append(line[9 * j:9 * j + 8].strip())
    t.tac = tac
    rec.skip()
    for i in range((t.natyps + 12) // 13 if t.natyps > 0 else 0):
        rec.skip()
    if t.version < 4:
        t.excl = np.zeros(t.nat_pro, np.int32)
        return
    # solvent type and boundary (L1041-1100)
    t.solvent_type = rec.values(1, int)[0]
    line = rec.line()
    first = line.split()[0] if line.split() else ""
    t.excl = np.zeros(t.nat_pro, np.int32)
    if first == "PBC":
        t.use_PBC = True
        t.boxlength = np.array(rec.values(3, _to_float))
        t.boxcentre = np.array(rec.values(3, _to_float))
    else:
        t.use_PBC = False
        v = rec.leading(line, _to_float, 3)
        if t.version < 5.01:
            t.rexcl_o, t.rwat = v[0], v[2]
        else:
            t.rexcl_o, t.rwat = v[0], v[1]
        t.xpcent = np.array(rec.values(3, _to_float))
        t.xwcent = np.array(rec.values(3, _to_float))
        rec.values(2, int)  # nexats, nexwat
        s = rec.fixed_chars(t.nat_pro)
        t.excl = (np.frombuffer(s.encode(), dtype=np.uint8) == ord("T")).astype(np.int32)
