"""Seeded synthetic systems of the shapes BASELINE.json names.

There is no network and ``/root/reference`` does not exist on the GPU box, so the
benchmark and the large parity runs use generated systems: rigid TIP3P waters on
a cubic lattice (a = 3.103 A, rho = 0.0335 A^-3, topo.f90:1068) with random
orientations, a "protein-like" core of neutral three-atom charge groups with
exclusions and 1-4 pairs, and a compact chain of Q-atoms in the centre.  Every
generator returns a :class:`QSystem` (tables laid out as the Fortran host holds
them) whose ``xtop`` are the coordinates.
"""
from __future__ import annotations

import numpy as np

from .system import QSystem, make_qconn
from .topo import MAX_NBR_RANGE

COULOMB = 332.0
A_LATTICE = 3.103
# (sqrt(A), sqrt(B)) normal and 1-4, OPLS-AA-like magnitudes (kcal/mol, A)
_TYPES = np.array([
    # mass   A1      A2    A3(1-4)  B1     B2    B3(1-4)
    [15.999, 762.89, 0.0, 539.45, 24.39, 0.0, 17.25],   # 1 water O (TIP3P)
    [1.008, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],              # 2 water H
    [12.011, 944.52, 0.0, 667.88, 22.03, 0.0, 15.58],   # 3 C sp3
    [1.008, 84.57, 0.0, 59.80, 5.41, 0.0, 3.83],        # 4 H
    [15.999, 616.44, 0.0, 435.89, 23.77, 0.0, 16.81],   # 5 O
    [14.007, 971.50, 0.0, 686.95, 26.15, 0.0, 18.49],   # 6 N
    [12.011, 1059.13, 0.0, 748.92, 23.67, 0.0, 16.74],  # 7 C aromatic
    [32.060, 2000.00, 0.0, 1414.2, 35.00, 0.0, 24.75],  # 8 S
    [1.008, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],              # 9 polar H
])
# the same atom types for the arithmetic rule (ivdw_rule 2): R* (half Rmin) in the A columns, epsilon in the B columns,
# as AMBER / CHARMM style libraries hold them (topo.f90:847-891; epsilon is square-rooted by topology(), simprep.f90:4567)
_TYPES_ARITH = np.array([
    # mass    R*1     R*2   R*3(1-4) eps1    eps2  eps3(1-4)
    [15.999, 1.7683, 0.0, 1.7683, 0.1520, 0.0, 0.0760],   # 1 water O
    [1.008, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],                # 2 water H (no LJ)
    [12.011, 1.9080, 0.0, 1.9080, 0.1094, 0.0, 0.0547],   # 3 C sp3
    [1.008, 1.4870, 0.0, 1.4870, 0.0157, 0.0, 0.0079],    # 4 H
    [15.999, 1.6612, 0.0, 1.6612, 0.2100, 0.0, 0.1050],   # 5 O
    [14.007, 1.8240, 0.0, 1.8240, 0.1700, 0.0, 0.0850],   # 6 N
    [12.011, 1.9080, 0.0, 1.9080, 0.0860, 0.0, 0.0430],   # 7 C aromatic
    [32.060, 2.0000, 0.0, 2.0000, 0.2500, 0.0, 0.1250],   # 8 S
    [1.008, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],                # 9 polar H
])
R_OH, ANG_HOH = 0.9572, np.deg2rad(104.52)
Q_O, Q_H = -0.834, 0.417


def _rand_rot(rng, n):
    """n uniform random rotation matrices (from unit quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)


def _waters(rng, centers):
    """Rigid TIP3P molecules at the given oxygen positions, random orientation: [n][3][3]."""
    n = len(centers)
    h1 = np.array([R_OH * np.sin(ANG_HOH / 2), 0.0, R_OH * np.cos(ANG_HOH / 2)])
    h2 = np.array([-R_OH * np.sin(ANG_HOH / 2), 0.0, R_OH * np.cos(ANG_HOH / 2)])
    R = _rand_rot(rng, n)
    out = np.zeros((n, 3, 3))
    out[:, 0] = centers
    out[:, 1] = centers + R @ h1
    out[:, 2] = centers + R @ h2
    return out


def _lattice(a, half):
    k = int(np.ceil(half / a)) + 1
    g = (np.arange(-k, k + 1) + 0.5) * a
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)


def _base(natyp_used=9, ivdw_rule=1, solvent_type=0) -> QSystem:
    q = QSystem()
    q.natyps = len(_TYPES)
    q.ivdw_rule = ivdw_rule
    q.types_file = (_TYPES if ivdw_rule == 1 else _TYPES_ARITH).copy()   # as a topology file holds them
    q.iaclib = q.types_file.copy()
    if ivdw_rule == 2:
        q.iaclib[:, 4:7] = np.sqrt(np.abs(q.iaclib[:, 4:7]))             # topology(), simprep.f90:4567-4571
    q.solvent_type = solvent_type
    q.el14_scale = 0.5
    q.iuse_switch_atom = 1
    return q


def _finish(q: QSystem, x, iac, crg, groups, nat_solute, nwat, bonds, listex_pairs, list14_pairs, q_atoms,
            nstates, qcrg=None, fep_types=None, softcore_alpha=None, soft_pairs=(), qbnd=None, el_scale=(),
            qq_use_library_charges=False):
    """Assemble the Fortran-layout tables. groups: list of (switch, [atoms]) 1-based."""
    # what a topology / FEP file of this system holds (write_files)
    q.raw = dict(crg=[float(c) for c in crg], bonds=[tuple(int(v) for v in b) for b in bonds],
                 qcrg=None if qcrg is None else np.asarray(qcrg, np.float64).copy(),
                 fep_types=fep_types, softcore_alpha=softcore_alpha, soft_pairs=list(soft_pairs), qbnd=qbnd,
                 el_scale=list(el_scale), types_file=q.types_file)
    q.qq_use_library_charges = int(qq_use_library_charges)
    q.natom, q.nat_solute, q.nwat, q.solv_atom = len(iac), nat_solute, nwat, 3
    q.iac = np.asarray(iac, np.int32)
    q.num_atyp = int(q.iac.max())
    q.ljcod = np.ones((q.num_atyp, q.num_atyp), np.int32)
    q.crg = np.asarray(crg, np.float64) * np.sqrt(COULOMB)
    q.xtop = np.asarray(x, np.float64).reshape(-1, 3)
    q.ncgp = len(groups)
    q.ncgp_solute = sum(1 for sw, at in groups if sw <= nat_solute)
    cgp = np.zeros((q.ncgp, 3), np.int32)
    cgpatom = []
    k = 1
    for g, (sw, atoms) in enumerate(groups):
        cgp[g] = (sw, k, k + len(atoms) - 1)
        cgpatom.extend(atoms)
        k += len(atoms)
    q.cgp, q.cgpatom = cgp, np.asarray(cgpatom, np.int32)
    q.excl = np.zeros(q.natom, np.int32)
    listex = np.zeros((nat_solute, MAX_NBR_RANGE), np.int32)
    list14 = np.zeros((nat_solute, MAX_NBR_RANGE), np.int32)
    exlong, l14long = [], []
    for (i, j) in listex_pairs:
        i, j = min(i, j), max(i, j)
        if j - i <= MAX_NBR_RANGE:
            listex[i - 1, j - i - 1] = 1
        else:
            exlong.append((i, j))
    for (i, j) in list14_pairs:
        i, j = min(i, j), max(i, j)
        if j - i <= MAX_NBR_RANGE:
            if not listex[i - 1, j - i - 1]:
                list14[i - 1, j - i - 1] = 1
        else:
            l14long.append((i, j))
    q.listex, q.list14 = listex, list14
    q.listexlong = np.asarray(exlong, np.int32).reshape(-1, 2)
    q.list14long = np.asarray(l14long, np.int32).reshape(-1, 2)
    # Q-atoms
    q.nqat, q.nstates = len(q_atoms), nstates
    q.iqseq = np.asarray(q_atoms, np.int32)
    q.iqatom = np.zeros(q.natom, np.int32)
    for k, a in enumerate(q_atoms):
        q.iqatom[a - 1] = k + 1
    nq = q.nqat
    if nq:
        base = np.asarray(crg, np.float64)[q.iqseq - 1]
        q.qcrg = (np.repeat(base[:, None], nstates, 1) if qcrg is None else np.asarray(qcrg, np.float64)) * np.sqrt(COULOMB)
        if fep_types is None:
            # no [change_atoms]: vdw_from_topo (qatom.f90:809-823)
            q.qvdw_flag, q.nqlib = 0, nq
            q.qavdw = q.iaclib[q.iac[q.iqseq - 1] - 1, 1:4].copy()
            q.qbvdw = q.iaclib[q.iac[q.iqseq - 1] - 1, 4:7].copy()
            q.qiac = np.repeat(np.arange(1, nq + 1, dtype=np.int32)[:, None], nstates, 1)
        else:
            lib_a, lib_b, qiac = fep_types
            q.qvdw_flag, q.nqlib = 1, len(lib_a)
            q.qavdw, q.qbvdw, q.qiac = np.array(lib_a, float), np.array(lib_b, float), np.asarray(qiac, np.int32)
            if q.ivdw_rule == 2:                 # get_fep, simprep.f90:1041-1046: epsilon of codes 1 and 3, not the soft pair's ai
                q.qbvdw[:, 0] = np.sqrt(q.qbvdw[:, 0])
                q.qbvdw[:, 2] = np.sqrt(q.qbvdw[:, 2])
        q.sc_lookup = np.zeros((nq, q.natyps + nq, nstates))
        if softcore_alpha is not None:
            al = np.asarray(softcore_alpha, float)      # [nq][nstates], plain alphas (qatom.f90:1932,1974)
            q.sc_lookup[:, :q.natyps, :] = al[:, None, :]
            q.sc_lookup[:, q.natyps:, :] = np.maximum(al[:, None, :], al[None, :, :])
            both0 = (al[:, None, :] <= 1e-6) & (al[None, :, :] <= 1e-6)
            q.sc_lookup[:, q.natyps:, :][both0] = 0.0
        q.iqexpnb = np.asarray([p[0] for p in soft_pairs], np.int32)
        q.jqexpnb = np.asarray([p[1] for p in soft_pairs], np.int32)
        q.el_scale_iq = np.asarray([e[0] for e in el_scale], np.int32)
        q.el_scale_jq = np.asarray([e[1] for e in el_scale], np.int32)
        q.el_scale = np.asarray([e[2] for e in el_scale], np.float64).reshape(len(el_scale), nstates)
        bnd = np.asarray(bonds, np.int32).reshape(-1, 3)
        qb_ij = np.zeros((0, 2), np.int32) if qbnd is None else np.asarray(qbnd[0], np.int32)
        qb_cod = np.zeros((0, nstates), np.int32) if qbnd is None else np.asarray(qbnd[1], np.int32)
        q.qconn = make_qconn(nstates, nat_solute, nq, q.iqseq, q.iqatom, bnd, qb_ij, qb_cod,
                             np.zeros((0, 2), np.int32), np.zeros((0, nstates), np.int32))
    else:
        q.qcrg = np.zeros((0, nstates))
        q.qiac = np.zeros((0, nstates), np.int32)
        q.qavdw = np.zeros((0, 3))
        q.qbvdw = np.zeros((0, 3))
        q.nqlib = 0
        q.sc_lookup = np.zeros((0, q.natyps, nstates))
        q.qconn = np.zeros(0, np.int32)
        q.iqexpnb = q.jqexpnb = q.el_scale_iq = q.el_scale_jq = np.zeros(0, np.int32)
        q.el_scale = np.zeros((0, nstates))
    return q.full_shard()


def water_box(n_side: int = 32, seed: int = 20261018, jitter: float = 0.05) -> QSystem:
    """C5: n^3 TIP3P waters in a periodic cube of edge n*3.103 A (n=32: 98 304 atoms, 99.3 A)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = n_side * A_LATTICE
    g = (np.arange(n_side) + 0.5) * A_LATTICE - L / 2
    centers = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    w = _waters(rng, centers)
    x = w.reshape(-1, 3)
    x = x + np.random.Generator(np.random.PCG64(seed + 1)).normal(0, jitter, x.shape)
    nw = len(centers)
    q = _base()
    q.use_PBC, q.use_LRF = 1, 1
    q.boxlength = np.array([L, L, L])
    iac = np.tile([1, 2, 2], nw)
    crg = np.tile([Q_O, Q_H, Q_H], nw)
    groups = [(3 * k + 1, [3 * k + 1, 3 * k + 2, 3 * k + 3]) for k in range(nw)]
    return _finish(q, x, iac, crg, groups, 0, nw, [], [], [], [], 1)


def _q_chain(rng, nq, center):
    """nq chain atoms: all-trans zig-zag strands of eight atoms (1.53 A bonds, 2.5 A 1-3 distance) laid
    side by side 4 A apart, so that atoms more than three bonds apart never overlap."""
    per, dx, dy, gap = 8, 1.25, 0.44, 4.0
    pts = []
    for k in range(nq):
        s, m = divmod(k, per)
        col, lay = s % 3, s // 3
        mm = m if s % 2 == 0 else per - 1 - m          # snake: strand ends stay adjacent
        pts.append((mm * dx, col * gap + (dy if m % 2 else -dy), lay * gap))
    pts = np.array(pts, float).reshape(-1, 3)
    if len(pts):
        pts -= (pts.max(0) + pts.min(0)) / 2
    return pts + center + rng.normal(0, 0.03, pts.shape)


def solvated_sphere(radius: float = 30.0, core_radius: float = 19.5, nq: int = 46, nstates: int = 1,
                    seed: int = 20261017, jitter: float = 0.05, fep: str = "none", pbc_box: float = 0.0,
                    excl_shell: float = 0.0, ivdw_rule: int = 1, solvent_type: int = 0, el_scale: bool = False,
                    qq_use_library_charges: bool = False) -> QSystem:
    """C2 (radius 30, 1 state) / C3 (radius 25, core 0, fep='annihilate', 2 states) / EVB-like ('evb').

    pbc_box > 0 builds the same content in a periodic cube instead of a sphere (tests of the box paths).
    excl_shell > 0 flags atoms further than radius-excl_shell from the centre as excluded.
    ivdw_rule 2 = arithmetic combination rule (R*, epsilon libraries); solvent_type 1 = a three-site solvent whose
    hydrogens carry LJ parameters (with either, potene.f90:347 takes the general ww/qw routines instead of the SPC ones);
    el_scale adds two [el_scale] pairs, qq_use_library_charges the [FEP] switch of that name.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    center = np.zeros(3)
    x, iac, crg, groups, bonds, ex, l14 = [], [], [], [], [], [], []
    # --- Q-atom chain (ligand), charge groups of up to four atoms
    qpos = _q_chain(rng, nq, center)
    q_extent = np.abs(qpos - center).max(0) + 2.2 if nq else np.zeros(3)
    qtypes = rng.choice([3, 4, 5, 6, 7], size=nq, p=[0.35, 0.35, 0.1, 0.1, 0.1])
    for k in range(nq):
        x.append(qpos[k]); iac.append(int(qtypes[k]))
    k = 0
    while k < nq:
        n = min(4, nq - k)
        ch = rng.uniform(-0.4, 0.4, n)
        ch -= ch.mean()
        crg.extend(ch.tolist())
        groups.append((k + 1, list(range(k + 1, k + n + 1))))
        k += n
    for k in range(1, nq):
        bonds.append((k, k + 1, 1))
    # --- protein-like core: neutral three-atom groups on a jittered 3.2 A lattice
    nat = nq
    if core_radius > 0:
        lat = _lattice(3.2, core_radius)
        r = np.linalg.norm(lat, axis=1)
        inbox = np.all(np.abs(lat) > 0, axis=1)
        keep = (r <= core_radius) & inbox & ~np.all(np.abs(lat) < q_extent + 1.2, axis=1)
        lat = lat[keep] + rng.uniform(-0.15, 0.15, (keep.sum(), 3))
        R = _rand_rot(rng, len(lat))
        a1 = np.array([0.9, 0.0, 0.0])
        a2 = np.array([0.9 * np.cos(np.deg2rad(110)), 0.9 * np.sin(np.deg2rad(110)), 0.0])
        first_atoms = []
        for g in range(len(lat)):
            c = lat[g]
            t0 = int(rng.choice([3, 5, 6, 7, 8], p=[0.5, 0.15, 0.15, 0.15, 0.05]))
            qh = rng.uniform(0.1, 0.4)
            x.extend([c, c + R[g] @ a1, c + R[g] @ a2])
            iac.extend([t0, 9, 9])      # polar hydrogens: charge only, no LJ (they may sit close)
            crg.extend([-2 * qh, qh, qh])
            i0 = nat + 1
            groups.append((i0, [i0, i0 + 1, i0 + 2]))
            bonds.extend([(i0, i0 + 1, 1), (i0, i0 + 2, 1)])
            ex.extend([(i0, i0 + 1), (i0, i0 + 2), (i0 + 1, i0 + 2)])
            first_atoms.append(i0)
            nat += 3
        for g in range(len(first_atoms) - 1):
            l14.append((first_atoms[g], first_atoms[g + 1]))          # index distance 3
            if g % 7 == 0:
                l14.append((first_atoms[g] + 1, first_atoms[g + 1] + 1))
        for g in range(0, len(first_atoms) - 12, 50):
            l14.append((first_atoms[g], first_atoms[g + 10]))         # index distance 30 > 25: long list
            ex.append((first_atoms[g] + 1, first_atoms[g + 12]))      # long exclusion
    nat_solute = nat
    # --- waters
    half = pbc_box / 2 if pbc_box > 0 else radius
    lat = _lattice(A_LATTICE, half)
    r = np.linalg.norm(lat, axis=1)
    if pbc_box > 0:
        n_side = int(round(pbc_box / A_LATTICE))
        g1 = (np.arange(n_side) + 0.5) * A_LATTICE - pbc_box / 2
        lat = np.stack(np.meshgrid(g1, g1, g1, indexing="ij"), -1).reshape(-1, 3)
        r = np.linalg.norm(lat, axis=1)
        keep = np.ones(len(lat), bool)
    else:
        keep = r <= radius
    if core_radius > 0:
        keep &= r > core_radius + 2.6
    else:
        keep &= ~np.all(np.abs(lat) < q_extent + 1.0, axis=1)
    w = _waters(rng, lat[keep])
    nwat = len(w)
    hw = 2 if solvent_type == 0 else 4      # general three-site solvent: hydrogens of a type with LJ parameters
    for k in range(nwat):
        i0 = nat + 1
        x.extend([w[k, 0], w[k, 1], w[k, 2]])
        iac.extend([1, hw, hw])
        crg.extend([Q_O, Q_H, Q_H])
        groups.append((i0, [i0, i0 + 1, i0 + 2]))
        nat += 3
    x = np.array(x)
    x = x + np.random.Generator(np.random.PCG64(seed + 1)).normal(0, jitter, x.shape)

    q = _base(ivdw_rule=ivdw_rule, solvent_type=solvent_type)
    q.use_LRF = 1
    if pbc_box > 0:
        q.use_PBC = 1
        q.boxlength = np.array([pbc_box] * 3)
        q.qswitch = min(7, nq) if nq else 0
    else:
        q.use_PBC = 0
        q.xpcent = center.copy()
        q.rexcl_o = radius
    q_atoms = list(range(1, nq + 1))
    kw = {}
    if nq and fep in ("annihilate", "evb"):
        assert nstates == 2
        # Q-atom type library: one row per topology type (normal, soft-pair Ci/ai, 1-4) + a dummy
        lib_a = np.vstack([q.types_file[:, 1:4], np.zeros((1, 3))])
        lib_b = np.vstack([q.types_file[:, 4:7], np.zeros((1, 3))])
        lib_a[:, 1] = 90.0   # Ci of the exponential repulsion (soft pairs)
        lib_b[:, 1] = 1.8    # ai
        dummy = len(_TYPES) + 1
        base = np.array(crg[:nq])
        if fep == "annihilate":
            qiac = np.stack([qtypes, np.full(nq, dummy)], 1)
            qc = np.stack([base, np.zeros(nq)], 1)
            alpha = np.stack([np.zeros(nq), np.full(nq, 20.0)], 1)
            kw = dict(qcrg=qc, fep_types=(lib_a, lib_b, qiac), softcore_alpha=alpha)
        else:
            t2 = qtypes.copy()
            t2[::5] = 5
            ch2 = base + rng.uniform(-0.2, 0.2, nq)
            ch2 -= ch2.mean() - base.mean()
            qiac = np.stack([qtypes, t2], 1)
            alpha = np.zeros((nq, 2))
            alpha[:4, 1] = 10.0
            # a bond that exists only in state 1 and one only in state 2 (changes qconn per state)
            qb = (np.array([[2, 9], [3, 12]], np.int32), np.array([[1, 0], [0, 1]], np.int32))
            kw = dict(qcrg=np.stack([base, ch2], 1), fep_types=(lib_a, lib_b, qiac), softcore_alpha=alpha,
                      soft_pairs=[p for p in ((5, 20), (6, 30)) if p[1] <= nq], qbnd=qb)
    if nq >= 6 and el_scale:
        kw["el_scale"] = [(1, 3, [0.5, 0.8][:nstates]), (6, 2, [0.25, 1.0][:nstates])]
    qs = _finish(q, x, iac, crg, groups, nat_solute, nwat, bonds, ex, l14, q_atoms, nstates,
                 qq_use_library_charges=qq_use_library_charges, **kw)
    if excl_shell > 0 and not pbc_box:
        rr = np.linalg.norm(qs.xtop - center, axis=1)
        # whole charge groups are excluded through their switch atom; flag all atoms of such groups
        for g in range(qs.ncgp):
            sw = qs.cgp[g, 0]
            if rr[sw - 1] > radius - excl_shell:
                for a in qs.cgpatom[qs.cgp[g, 1] - 1:qs.cgp[g, 2]]:
                    qs.excl[a - 1] = 1
    return qs


def water_constraints(q: QSystem):
    """What init_constraints (simprep.f90:2167-2345) builds for a synthetic system with the default input (solvent bonds
    to hydrogens): rows (i, j, dist2) with the bond lengths write_files puts into the topology, the first atom of every
    molecule, and the inverse masses."""
    hh = (2.0 * R_OH * float(np.sin(ANG_HOH / 2))) ** 2
    cons = []
    for k in range(q.nwat):
        o = q.nat_solute + 3 * k + 1
        cons += [(o, o + 1, R_OH ** 2), (o, o + 2, R_OH ** 2), (o + 1, o + 2, hh)]
    starts = ([1] if q.nat_solute else []) + [q.nat_solute + 3 * k + 1 for k in range(q.nwat)]
    mass = np.asarray(q.iaclib).reshape(-1, 7)[np.asarray(q.iac) - 1, 0]
    return cons, starts, 1.0 / mass


def _wrap(items, per_line, fmt=str):
    items = [fmt(v) for v in items]
    return [" ".join(items[i:i + per_line]) for i in range(0, len(items), per_line)]


def write_files(q: QSystem, top_path: str, fep_path: str | None = None) -> None:
    """Writes a synthetic system as Qdyn6 input files -- a version-5 Q topology (the record order topo_read expects,
    topo.f90:522-1114) and, when it has Q-atoms, an FEP file (qatom.f90 sections) -- so that the unchanged readers
    (q6_b200/topo.py + fep.py, q6_b200/host/qdyn_host.cpp) and the compiled driver can be run where the reference's
    shipped inputs do not exist.  Numbers are written with repr(): reading the files back reproduces every table of q
    bit for bit (tests/test_host_cpp.py)."""
    raw = q.raw
    nat, nsol, nw = q.natom, q.nat_solute, q.nwat
    f = repr
    L = ["Q topology file", "TITLE      synthetic system (q6_b200.synth)", "VERSION     5.03", "END        of header",
         f"{nat} {nsol} = Total no. of atoms, no. of solute atoms. Coordinates: (2*3 per line)"]
    L += _wrap(np.asarray(q.xtop).reshape(-1).tolist(), 6, f)
    L.append(f"{nat} = No. of integer atom codes. Array:")
    L += _wrap(np.asarray(q.iac).tolist(), 20)
    # bonds: solute bonds (code 1), then O-H, O-H, H-H of every water (codes 2, 2, 3) -- the SHAKE constraints
    bonds = list(raw["bonds"])
    nb_sol = len(bonds)
    for k in range(nw):
        o = nsol + 3 * k + 1
        bonds += [(o, o + 1, 2), (o, o + 2, 2), (o + 1, o + 2, 3)]
    L.append(f"{len(bonds)} {nb_sol} = No. of bonds, no. of solute bonds. i - j - icode: (5 per line)")
    L += _wrap([v for b in bonds for v in b], 15)
    L.append("3 = No. of bond codes. Parameters: fk, r0")
    h_h = 2.0 * R_OH * float(np.sin(ANG_HOH / 2))
    L += [f"1 {f(300.0)} {f(1.53)}", f"2 {f(1106.0)} {f(R_OH)}", f"3 {f(0.0)} {f(h_h)}"]
    L += ["0 0 = No. of angles, no. of solute angles", "0 = No. of angle codes",
          "0 0 = No. of torsions, solute torsions", "0 = No. of torsion codes",
          "0 0 = No. of impropers, solute impropers", "0 = No. of improper codes",
          f"{nat} = No. of atomic charges"]
    L += _wrap(raw["crg"][:nat], 6, f)
    L.append(f"{q.ncgp} {q.ncgp_solute} {q.iuse_switch_atom} = No. of charge groups, no. of solute cgps, switch-atom flag")
    cgp = np.asarray(q.cgp).reshape(-1, 3)
    for g in range(q.ncgp):
        isw, first, last = (int(v) for v in cgp[g])
        L.append(f"{last - first + 1} {isw}")
        L += _wrap(np.asarray(q.cgpatom)[first - 1:last].tolist(), 20)
    lib = np.asarray(raw["types_file"]).reshape(-1, 7)      # epsilon not yet square-rooted for the arithmetic rule
    L += [f"{q.natyps} = No. of atom types", f"{q.ivdw_rule} = vdW combination rule",
          f"{f(float(q.el14_scale))} {f(COULOMB)} = Electrostatic 1-4 scaling factor and Coulomb constant", "Masses:"]
    L += _wrap(lib[:, 0].tolist(), 6, f)
    for j, nm in enumerate(("1", "2", "3")):
        L.append(f"Av/Aii({nm}) terms:")
        L += _wrap(lib[:, 1 + j].tolist(), 6, f)
        L.append(f"Bv/Bii({nm}) terms:")
        L += _wrap(lib[:, 4 + j].tolist(), 6, f)
    L.append("0 = No. of type-2 vdW interactions")

    def bit_rows(tab):
        flat = "".join("1" if v else "0" for v in np.asarray(tab).reshape(-1))
        return [flat[i:i + 80] for i in range(0, len(flat), 80)]

    l14long = np.asarray(q.list14long).reshape(-1, 2)
    exlong = np.asarray(q.listexlong).reshape(-1, 2)
    L.append(f"{int(np.asarray(q.list14).sum())} = No. of 1-4 neighbours. List (range={MAX_NBR_RANGE})")
    if nsol > 0:
        L += bit_rows(q.list14)
    L.append(f"{len(l14long)} = No. of long-range 1-4 nbrs (>{MAX_NBR_RANGE})")
    L += [f"{int(a)} {int(b)}" for a, b in l14long]
    L.append(f"{int(np.asarray(q.listex).sum())} = No. of exclusions. List (range={MAX_NBR_RANGE})")
    if nsol > 0:
        L += bit_rows(q.listex)
    L.append(f"{len(exlong)} = No. of long-range exclusions (>{MAX_NBR_RANGE})")
    L += [f"{int(a)} {int(b)}" for a, b in exlong]
    # one residue / molecule for the solute, one per water
    starts = ([1] if nsol > 0 else []) + [nsol + 3 * k + 1 for k in range(nw)]
    names = (["SYN "] if nsol > 0 else []) + ["HOH "] * nw
    L.append(f"{len(starts)} {1 if nsol > 0 else 0} = No. of residues, No of solute residues. start atoms:")
    L += _wrap(starts, 13)
    L.append("Sequence:")
    L += ["".join(n.ljust(5) for n in names[i:i + 16]) for i in range(0, len(names), 16)]
    L.append(f"{len(starts)} = No. of separate molecules. start atoms:")
    L += _wrap(starts, 13)
    L.append("Atom type names:")
    tac = [f"T{i + 1}" for i in range(q.natyps)]
    L += ["".join(n.ljust(9) for n in tac[i:i + 8]) for i in range(0, len(tac), 8)]
    L.append("SYBYL atom types:")
    L += ["".join("Du   " for _ in tac[i:i + 13]) for i in range(0, len(tac), 13)]
    L.append(f"{q.solvent_type} = solvent type (0=SPC,1=3-atom,2=general)")
    if q.use_PBC:
        L.append("PBC = boundary condition")
        L.append(" ".join(f(float(v)) for v in q.boxlength) + " = Boxlength")
        L.append("0.0 0.0 0.0 = Centre of the box")
    else:
        L.append(f"{f(float(q.rexcl_o))} {f(float(q.rexcl_o))} {f(float(q.rexcl_o))} = Exclusion, eff. solvent & solvent radius")
        L.append(" ".join(f(float(v)) for v in q.xpcent) + " = Solute centre")
        L.append(" ".join(f(float(v)) for v in q.xpcent) + " = Solvent centre")
        ex = np.asarray(q.excl).astype(bool)
        L.append(f"{int(ex.sum())} {int(ex[nsol:].sum()) // 3} = No. of excluded atoms (incl. water), no. of excluded waters")
        flat = "".join("T" if v else "F" for v in ex)
        L += [flat[i:i + 80] for i in range(0, len(flat), 80)]
    with open(top_path, "w") as fh:
        fh.write("\n".join(L) + "\n")
    if not q.nqat or fep_path is None:
        return
    ns, nq = q.nstates, q.nqat
    F = ["! synthetic FEP file (q6_b200.synth)", "[FEP]", f"states {ns}"]
    if q.qq_use_library_charges:
        F.append("qq_use_library_charges on")
    if q.use_PBC:
        F += ["[PBC]", f"switching_atom {q.qswitch}"]
    F.append("[atoms]")
    F += [f"{k + 1} {int(a)}" for k, a in enumerate(q.iqseq)]
    if raw["qcrg"] is not None:
        F.append("[change_charges]")
        F += [f"{k + 1} " + " ".join(f(float(v)) for v in raw["qcrg"][k]) for k in range(nq)]
    if raw["fep_types"] is not None:
        lib_a, lib_b, qiac = raw["fep_types"]
        F.append("[atom_types]")
        for i in range(len(lib_a)):
            ab = " ".join(f"{f(float(lib_a[i][k]))} {f(float(lib_b[i][k]))}" for k in range(3))
            F.append(f"Q{i + 1} {ab} {f(12.0)}")
        F.append("[change_atoms]")
        F += [f"{k + 1} " + " ".join(f"Q{int(t)}" for t in np.asarray(qiac)[k]) for k in range(nq)]
    if raw["soft_pairs"]:
        F.append("[soft_pairs]")
        F += [f"{int(a)} {int(b)}" for a, b in raw["soft_pairs"]]
    if raw["el_scale"]:
        F.append("[el_scale]")
        F += [f"{int(a)} {int(b)} " + " ".join(f(float(v)) for v in sc) for a, b, sc in raw["el_scale"]]
    if raw["qbnd"] is not None:
        F.append("[change_bonds]")
        F += [f"{int(i)} {int(j)} " + " ".join(str(int(c)) for c in cod) for (i, j), cod in zip(*raw["qbnd"])]
    if raw["softcore_alpha"] is not None:
        F.append("[softcore]")
        F += [f"{k + 1} " + " ".join(f(float(v)) for v in np.asarray(raw["softcore_alpha"])[k]) for k in range(nq)]
    with open(fep_path, "w") as fh:
        fh.write("\n".join(F) + "\n")


def config(name: str) -> tuple:
    """(QSystem, cut-offs dict, lambdas) of a named BASELINE.json configuration."""
    cuts_sph = dict(Rq=99.0, Rcq2=99.0 ** 2, RcLRF2=99.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=99.0)
    if name == "C2":
        return solvated_sphere(32.0, 19.5, 46, 1, 20261017), cuts_sph, np.array([1.0])
    if name == "C3":
        return solvated_sphere(25.0, 0.0, 46, 2, 20261019, fep="annihilate"), cuts_sph, np.array([0.5, 0.5])
    if name == "C4s":
        return solvated_sphere(25.0, 15.0, 60, 2, 20261020, fep="evb"), cuts_sph, np.array([0.5, 0.5])
    if name == "C5":
        cuts = dict(Rq=-1.0, Rcq2=1.0, RcLRF2=24.0 ** 2, Rcpp2=100.0, Rcpw2=100.0, Rcww2=100.0, RcLRF=24.0)
        return water_box(32, 20261018), cuts, np.array([1.0])
    raise KeyError(name)
