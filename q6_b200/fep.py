"""FEP/EVB strategy file reader: the sections that feed the nonbonded path.

Mirrors ``qatom_load_atoms`` (reference src/qatom.f90:322-640) and the
nonbonded-relevant part of ``qatom_load_fep`` (qatom.f90:639-1050, 1865-1996):
``[FEP]``, ``[PBC]``, ``[atoms]``, ``[change_charges]``, ``[atom_types]``,
``[change_atoms]``, ``[soft_pairs]``, ``[el_scale]``, ``[excluded_pairs]``,
``[softcore]`` and -- because ``make_qconn`` walks them -- ``[change_bonds]``.
The file syntax is prmfile.f90's (sections in brackets, ``!`` comments anywhere,
``#``/``*`` comments at line start or after white space).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .topo import NLJTYP, Topology


class FepError(RuntimeError):
    pass


def _strip_comment(line: str) -> str:
    # prmfile.f90:845-879
    out = line
    i = 0
    while i < len(out):
        c = out[i]
        if c == "!":
            return out[:i].rstrip()
        if c in "#*" and (i == 0 or out[i - 1] in " \t"):
            return out[:i].rstrip()
        i += 1
    return out.rstrip()


def parse_sections(text: str) -> dict:
    sections: dict = {}
    cur = None
    for raw in text.splitlines():
        s = raw.strip()
        if not s or s[0] in "!#*":
            continue
        if s.startswith("["):
            cur = s[1:s.index("]")].strip().lower()
            sections.setdefault(cur, [])
            continue
        s = _strip_comment(s)
        if s and cur is not None:
            sections[cur].append(s)
    return sections


def _logical(s: str) -> bool:
    return s.strip().lower() in ("on", "true", ".true.", "yes", "1", "t")


@dataclass
class Fep:
    nstates: int = 1
    nqat: int = 0
    offset: int = 0
    qq_use_library_charges: bool = False
    softcore_use_max_potential: bool = False
    qswitch: int = 0
    iqseq: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    qcrg: np.ndarray = field(default_factory=lambda: np.zeros((0, 1)))  # [nqat][nstates], unscaled
    qvdw_flag: bool = False
    nqlib: int = 0
    qtac: list = field(default_factory=list)
    qavdw: np.ndarray = field(default_factory=lambda: np.zeros((0, NLJTYP)))  # [nqlib][3]
    qbvdw: np.ndarray = field(default_factory=lambda: np.zeros((0, NLJTYP)))
    qiac: np.ndarray = field(default_factory=lambda: np.zeros((0, 1), np.int32))  # [nqat][nstates]
    iqexpnb: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    jqexpnb: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    el_scale_iq: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    el_scale_jq: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    el_scale: np.ndarray = field(default_factory=lambda: np.zeros((0, 1)))  # [n][nstates]
    exspec_ij: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    exspec_flag: np.ndarray = field(default_factory=lambda: np.zeros((0, 1), np.int32))  # [n][nstates]
    qbnd_ij: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    qbnd_cod: np.ndarray = field(default_factory=lambda: np.zeros((0, 1), np.int32))  # [n][nstates]
    alpha_max: np.ndarray = field(default_factory=lambda: np.zeros((0, 1)))  # [nqat][nstates]
    sc_lookup: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 1)))  # [nqat][natyps+nqat][nstates]


def load_fep(path: str, topo: Topology, nstates: int | None = None) -> Fep:
    """qatom_load_atoms + qatom_load_fep.  ``nstates`` is the [lambdas] count of the md input
    (checked against the file like qatom.f90:356-378); None = trust the file."""
    with open(path, "r", errors="replace") as fh:
        sec = parse_sections(fh.read())
    f = Fep()

    # ---- qatom_load_atoms (qatom.f90:322-640)
    if "atoms" not in sec:
        raise FepError(">>> WARNING: No [atoms] section in fep file. Aborting file loading.")
    fepsec = dict()
    for ln in sec.get("fep", []):
        parts = ln.split(None, 1)
        fepsec[parts[0].lower()] = parts[1].strip() if len(parts) > 1 else ""
    f.nstates = int(fepsec["states"]) if "states" in fepsec else 1
    if nstates is not None and f.nstates != nstates:
        raise FepError(">>>>> ERROR: Mismatch between nstates in input and FEP file")
    if f.nstates == 0:
        raise FepError(">>>>> ERROR: Number of states must be at least 1. Aborting.")
    f.qq_use_library_charges = _logical(fepsec.get("qq_use_library_charges", "off"))
    f.softcore_use_max_potential = _logical(fepsec.get("softcore_use_max_potential", "off"))
    offset = 0
    if "offset" in fepsec:
        offset = int(fepsec["offset"])
        if offset < 1 or offset > topo.nat_solute:
            raise FepError(f">>>>> ERROR: Invalid topology atom number offset value:{offset}")
    elif "offset_name" in fepsec:
        name = fepsec["offset_name"][:4]
        offset = -1
        for i in range(topo.nres_solute):
            if topo.res_name[i].strip() == name.strip():
                offset = int(topo.res_start[i]) - 1
                break
        if offset == -1:
            raise FepError(f">>>>> ERROR: Residue name {name} not found.")
    elif "offset_residue" in fepsec:
        r = int(fepsec["offset_residue"])
        if r < 1 or r > topo.nres_solute:
            raise FepError(f">>>>> ERROR: Invalid residue number for offset:{r}")
        offset = int(topo.res_start[r - 1]) - 1
    f.offset = offset

    rows = [ln.split() for ln in sec["atoms"]]
    if not rows:
        return f  # zero Q-atoms: fep file is not loaded (qatom.f90:434-439)
    try:
        pairs = [(int(r[0]), int(r[1])) for r in rows]
    except ValueError as e:
        raise FepError("only the 'int int' form of [atoms] is supported by this reader") from e
    f.nqat = max(p[0] for p in pairs)
    f.iqseq = np.zeros(f.nqat, np.int32)
    for s, topno in pairs:
        if topno + offset < 1 or topno + offset > topo.nat_solute:
            raise FepError(f">>>>> ERROR: invalid topology atom number {topno + offset} for Q-atom {s}")
        f.iqseq[s - 1] = topno + offset
    nq, ns = f.nqat, f.nstates

    # ---- get_fep copies the topology charges first (simprep.f90:1027-1031)
    f.qcrg = np.repeat(topo.crg[f.iqseq - 1][:, None], ns, axis=1).astype(np.float64)

    # ---- qatom_load_fep
    if topo.use_PBC:
        if "pbc" not in sec:
            raise FepError(">>>>> ERROR: Section PBC is required when using periodic boundary.")
        kv = dict(ln.split(None, 1) for ln in sec["pbc"])
        if "switching_atom" not in kv:
            raise FepError(">>>>> ERROR: Switching atom could not be read.")
        f.qswitch = int(kv["switching_atom"]) + offset

    for ln in sec.get("change_charges", []):  # qatom.f90:731-764
        tok = ln.split()
        iat = int(tok[0])
        if iat < 1 or iat > nq or f.iqseq[iat - 1] == 0:
            raise FepError(f">>>>> ERROR: {iat} is not a valid q-atom number")
        f.qcrg[iat - 1, :] = [float(v) for v in tok[1:1 + ns]]

    types = sec.get("atom_types", [])  # qatom.f90:780-823
    vdw_from_topo = False
    index = {}
    if types:
        f.nqlib = len(types)
        f.qavdw = np.zeros((f.nqlib, NLJTYP))
        f.qbvdw = np.zeros((f.nqlib, NLJTYP))
        for i, ln in enumerate(types):
            tok = ln.split()
            f.qtac.append(tok[0])
            vals = [float(v) for v in tok[1:1 + 2 * NLJTYP + 1]]
            for k in range(NLJTYP):
                f.qavdw[i, k] = vals[2 * k]
                f.qbvdw[i, k] = vals[2 * k + 1]
            if tok[0] in index:
                raise FepError(f">>>>> ERROR: Could not enumerate q-atom type {tok[0]} Duplicate name?")
            index[tok[0]] = i + 1
    else:
        vdw_from_topo = True
        f.nqlib = nq
        f.qavdw = np.zeros((nq, NLJTYP))
        f.qbvdw = np.zeros((nq, NLJTYP))
        for i in range(nq):
            t = topo.iac[f.iqseq[i] - 1]
            f.qtac.append(topo.tac[t - 1] if topo.tac else str(t))
            f.qavdw[i, :] = topo.iaclib[t - 1, 1:4]
            f.qbvdw[i, :] = topo.iaclib[t - 1, 4:7]

    f.qiac = np.zeros((nq, ns), np.int32)
    chg = sec.get("change_atoms", [])  # qatom.f90:825-866
    if not chg:
        f.qvdw_flag = False
        if vdw_from_topo:
            f.qiac[:, :] = np.arange(1, nq + 1, dtype=np.int32)[:, None]
    else:
        iat = max(int(ln.split()[0]) for ln in chg)
        if iat != nq or len(chg) != nq:
            raise FepError(">>>>> ERROR: Atom types of Q-atoms must be given for every Q-atom!")
        f.qvdw_flag = True
        for ln in chg:
            tok = ln.split()
            iat = int(tok[0])
            for j in range(ns):
                if tok[1 + j] not in index:
                    raise FepError(f">>>>> ERROR: Q-atom type {tok[1 + j]} has not been defined.")
                f.qiac[iat - 1, j] = index[tok[1 + j]]

    sp = sec.get("soft_pairs", [])  # qatom.f90:872-897
    f.iqexpnb = np.zeros(len(sp), np.int32)
    f.jqexpnb = np.zeros(len(sp), np.int32)
    for i, ln in enumerate(sp):
        j, k = (int(v) for v in ln.split()[:2])
        if j < 1 or j > nq or k < 1 or k > nq or f.iqseq[j - 1] == 0 or f.iqseq[k - 1] == 0:
            raise FepError(f">>>>> ERROR: Invalid q-atom number in this group: {j} {k}")
        f.iqexpnb[i], f.jqexpnb[i] = j, k

    es = sec.get("el_scale", [])  # qatom.f90:903-930
    f.el_scale_iq = np.zeros(len(es), np.int32)
    f.el_scale_jq = np.zeros(len(es), np.int32)
    f.el_scale = np.zeros((len(es), ns))
    for i, ln in enumerate(es):
        tok = ln.split()
        j, k = int(tok[0]), int(tok[1])
        if j < 1 or j > nq or k < 1 or k > nq or f.iqseq[j - 1] == 0 or f.iqseq[k - 1] == 0:
            raise FepError(f">>>>> ERROR: Invalid q-atom number in this group: {j} {k}")
        f.el_scale_iq[i], f.el_scale_jq[i] = j, k
        f.el_scale[i, :] = [float(v) for v in tok[2:2 + ns]]

    ex = sec.get("excluded_pairs", [])  # qatom.f90:936-965
    f.exspec_ij = np.zeros((len(ex), 2), np.int32)
    f.exspec_flag = np.zeros((len(ex), ns), np.int32)
    for i, ln in enumerate(ex):
        tok = ln.split()
        f.exspec_ij[i] = (int(tok[0]) + offset, int(tok[1]) + offset)
        for j in range(ns):
            v = int(tok[2 + j])
            if v not in (0, 1):
                raise FepError(">>>>> ERROR: Special exclusion state flags are invalid.")
            f.exspec_flag[i, j] = v

    cb = sec.get("change_bonds", [])  # qatom.f90:990-1018
    f.qbnd_ij = np.zeros((len(cb), 2), np.int32)
    f.qbnd_cod = np.zeros((len(cb), ns), np.int32)
    for i, ln in enumerate(cb):
        tok = ln.split()
        f.qbnd_ij[i] = (int(tok[0]) + offset, int(tok[1]) + offset)
        f.qbnd_cod[i, :] = [int(v) for v in tok[2:2 + ns]]

    _softcore(sec, f, topo)
    return f


def _softcore(sec: dict, f: Fep, topo: Topology) -> None:
    """[softcore] and the sc_lookup table (qatom.f90:1865-1991)."""
    nq, ns, nt = f.nqat, f.nstates, topo.natyps
    f.sc_lookup = np.zeros((nq, nt + nq, ns))
    f.alpha_max = np.zeros((nq, ns))
    if "softcore" not in sec:
        return
    if not f.qvdw_flag:
        raise FepError('>>>>> ERROR: Q-atom types must be redefined in "change_atoms" section')
    lines = sec["softcore"]
    if len(lines) != nq:
        raise FepError(">>>>> ERROR: Alpha must be given for every Q-atom!")
    for ln in lines:
        tok = ln.split()
        f.alpha_max[int(tok[0]) - 1, :] = [float(v) for v in tok[1:1 + ns]]
    geom = topo.ivdw_rule == 1
    am = f.alpha_max
    for i in range(nq):
        for s in range(ns):
            aq = f.qavdw[f.qiac[i, s] - 1, 0]
            bq = f.qbvdw[f.qiac[i, s] - 1, 0]
            for j in range(nt):  # q - surroundings
                if f.softcore_use_max_potential:
                    aj, bj = topo.iaclib[j, 1], topo.iaclib[j, 4]
                    if not geom:
                        # qdyn.f90:114-116: topology() has already square-rooted the library epsilons
                        # (simprep.f90:4567-4571) when get_fep -> qatom_load_fep runs; the Q-atom epsilons have not
                        bj = np.sqrt(abs(bj))
                    if am[i, s] > 1e-6:
                        if geom:
                            f.sc_lookup[i, j, s] = (-bq * bj + np.sqrt(bq * bq * bj * bj + 4.0 * am[i, s] * aq * aj)) / (
                                2.0 * am[i, s])
                        else:
                            f.sc_lookup[i, j, s] = (-2.0 * np.sqrt(bq) * bj + 2.0 * np.sqrt(
                                bq * bj ** 2 + am[i, s] * np.sqrt(bq) * bj)) * (aq + aj) ** 6 / (2.0 * am[i, s])
                else:
                    f.sc_lookup[i, j, s] = am[i, s]
            for j in range(nq):  # q - q
                if am[i, s] > 1e-6 or am[j, s] > 1e-6:
                    aj = f.qavdw[f.qiac[j, s] - 1, 0]
                    bj = f.qbvdw[f.qiac[j, s] - 1, 0]
                    if f.softcore_use_max_potential:
                        a = min(am[i, s], am[j, s])
                        if am[i, s] < 1e-6 or am[j, s] < 1e-6:
                            a = max(am[i, s], am[j, s])
                        if geom:
                            v = (-bq * bj + np.sqrt(bq * bq * bj * bj + 4.0 * a * aq * aj)) / (2.0 * a)
                        else:
                            v = (-2.0 * np.sqrt(bq * bj) + 2.0 * np.sqrt(bq * bj + a * np.sqrt(bq * bj))) * (
                                aq + aj) ** 6 / (2.0 * a)
                        f.sc_lookup[i, nt + j, s] = v
                    else:
                        f.sc_lookup[i, nt + j, s] = max(am[i, s], am[j, s])
