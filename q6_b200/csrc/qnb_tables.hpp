// qnb_tables.hpp -- host-side construction of the static tables the kernels read.
//
// This is the library's own statement of what the reference builds once in
// precompute_interactions (simprep.f90:2860-3589) and make_nbqqlist / nbqqlist
// (nonbondene.f90:729-745, 3231-3282).  The reference materialises pp_map
// (nat_solute^2 int32) and pp_precomp; here solute-solute pairs are described by
// per-atom LJ/charge data plus a sparse "special pair" table (excluded and 1-4
// pairs), and the Q-atom tables are dense [iq][state][atom] arrays laid out for
// coalesced device reads.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qnb.h"

namespace qnb {

constexpr double kSetEps = 1e-10;   // QREAL_EPS, sizes.f90:50
constexpr int kMaxStates = 8;

struct QPar { double A, B, el, score; };  // vdWA, vdWB, elec, score

struct StaticQPair {   // one entry of nbqq / nbqqp for one state
    int i, j;          // topology atoms (0-based): Q-atom first
    int iq, jq;        // Q-atom numbers (1-based) / for qqp: jq = 0
    int state;         // 0-based
    int soft;
    QPar p;
};

// pair codes of the sparse special table
enum : uint8_t { kPairExcluded = 0, kPair14 = 3 };

struct HostTables {
    qnb_system s{};      // scalars (pointers are NOT valid after init)
    // --- atoms (0-based indices everywhere below)
    std::vector<int> iac0;            // type index 0-based (full type space)
    std::vector<int> ctype;           // compact type per atom
    std::vector<int> ct_to_iac;       // compact type -> iac (1-based)
    int nct = 0;
    std::vector<double> lj_a, lj_b;   // [nct][3] avdw/bvdw by LJ code 1..3
    std::vector<uint8_t> ljcode;      // [nct][nct] 1 or 2
    std::vector<double> crg;
    std::vector<uint8_t> is_q, excl, qbonded;  // qbonded: any(qconn(:,i,:) <= 3)
    std::vector<int> grp_of_atom;     // iwhich_cgp - 1
    // --- charge groups
    std::vector<int> g_first, g_n, g_switch, g_atoms;
    // --- units: solute groups then waters
    int nunit = 0;
    std::vector<int> u_sw, u_grp;
    std::vector<uint8_t> u_excl;
    // --- sparse special solute-solute pairs, CSR per atom, sorted by partner
    std::vector<int> sp_off, sp_partner;
    std::vector<uint8_t> sp_code;
    // per group: sorted atoms that are special to at least one atom of the group (incl. its own atoms)
    std::vector<int> gs_off, gs_atoms;
    std::vector<int> nq_off, nq_atoms;   // non-Q atoms of every solute group, group order (tiles of the solute kernel)
    // --- water site parameters
    std::vector<double> w_crg;            // [solv_atom]
    std::vector<int> w_ctype;             // [solv_atom]
    std::vector<QPar> ww_par;             // [solv_atom][solv_atom]
    // --- Q tables
    std::vector<int> iqseq0;              // topology atom per Q-atom (0-based)
    std::vector<QPar> qp_tab;             // [(iq*nstates+s)][nat_solute]
    std::vector<uint8_t> qp_set;          // same shape
    std::vector<QPar> qw_tab;             // [(iq*nstates+s)][solv_atom]
    std::vector<StaticQPair> qq_list, qqp_list;
    std::string error;

    // ---------------- helpers
    double avdw(int iac1, int code) const { return s_iaclib[7 * (iac1 - 1) + 1 + (code - 1)]; }
    double bvdw(int iac1, int code) const { return s_iaclib[7 * (iac1 - 1) + 4 + (code - 1)]; }
    int ljcod(int a1, int b1) const { return s_ljcod[(a1 - 1) + (size_t)(b1 - 1) * s.num_atyp]; }
    int qconn(int is1, int i1, int iq1) const {
        return s_qconn[(is1 - 1) + (size_t)(i1 - 1) * s.nstates + (size_t)(iq1 - 1) * s.nstates * s.nat_solute];
    }
    double qcrg(int iq1, int is1) const { return s_qcrg[(iq1 - 1) + (size_t)(is1 - 1) * s.nqat]; }
    int qiac(int iq1, int is1) const { return s_qiac[(iq1 - 1) + (size_t)(is1 - 1) * s.nqat]; }
    double qavdw(int t1, int c) const { return s_qavdw[(t1 - 1) + (size_t)(c - 1) * s.nqlib]; }
    double qbvdw(int t1, int c) const { return s_qbvdw[(t1 - 1) + (size_t)(c - 1) * s.nqlib]; }
    double sc_lookup(int iq1, int k1, int is1) const {
        return s_sc[(iq1 - 1) + (size_t)(k1 - 1) * s.nqat + (size_t)(is1 - 1) * s.nqat * (s.natyps + s.nqat)];
    }

    // combination rule on two (a,b) parameter pairs: precompute_set_values_* (simprep.f90:3326-3342 etc.)
    void combine(double ai, double bi, double aj, double bj, double &A, double &B) const {
        if (s.ivdw_rule == QNB_VDW_GEOMETRIC) {
            A = ai * aj;
            B = bi * bj;
        } else {
            double t = ai + aj;
            t = t * t;
            t = t * t * t;
            double e = bi * bj;
            A = (t * t) * e;
            B = 2.0 * t * e;
        }
    }
    static bool is_set(const QPar &p) {
        return std::fabs(p.A) > kSetEps || std::fabs(p.B) > kSetEps || std::fabs(p.el) > kSetEps;
    }

    // Classification of a solute-solute atom pair (0-based, both non-Q): what pp_int_comp
    // (simprep.f90:2952-3065) records for it.  Returns false for an excluded pair.
    bool pp_code(int i, int j, int &code) const {
        const int *b = sp_partner.data() + sp_off[i], *e = sp_partner.data() + sp_off[i + 1];
        const int *it = std::lower_bound(b, e, j);
        if (it != e && *it == j) {
            uint8_t c = sp_code[it - sp_partner.data()];
            if (c == kPairExcluded) return false;
            code = 3;
            return true;
        }
        code = ljcode[ctype[i] * nct + ctype[j]];
        return true;
    }
    // Full pp parameters of a listed pair; set=false when the reference drops it (L1612-1613).
    bool pp_params(int i, int j, QPar &p, bool &set) const {
        int code;
        if (!pp_code(i, j, code)) return false;
        combine(lj_a[ctype[i] * 3 + code - 1], lj_b[ctype[i] * 3 + code - 1], lj_a[ctype[j] * 3 + code - 1],
                lj_b[ctype[j] * 3 + code - 1], p.A, p.B);
        p.el = crg[i] * crg[j];
        p.score = 0;
        set = is_set(p);
        if (code == 3) p.el *= s.el14_scale;
        return true;
    }
    // pw_precomp(i, site) (simprep.f90:3355-3384)
    QPar pw_params(int i, int site) const {
        QPar p{};
        int code = ljcode[ctype[i] * nct + w_ctype[site]];
        combine(lj_a[ctype[i] * 3 + code - 1], lj_b[ctype[i] * 3 + code - 1], lj_a[w_ctype[site] * 3 + code - 1],
                lj_b[w_ctype[site] * 3 + code - 1], p.A, p.B);
        p.el = crg[i] * w_crg[site];
        return p;
    }

    bool build(const qnb_system *sys);

  private:
    // copies of the caller's tables, valid only during build()
    const double *s_iaclib = nullptr, *s_qcrg = nullptr, *s_qavdw = nullptr, *s_qbvdw = nullptr, *s_sc = nullptr;
    const int32_t *s_ljcod = nullptr, *s_qconn = nullptr, *s_qiac = nullptr;
    QPar qp_values(int iq1, int j1, int is1, int vdw, bool &set) const;
    void build_specials(const qnb_system *sys);
    void build_q(const qnb_system *sys);
};

inline void HostTables::build_specials(const qnb_system *sys) {
    const int ns = s.nat_solute, R = s.max_nbr_range;
    std::vector<std::vector<std::pair<int, uint8_t>>> sp(ns);
    auto add = [&](int i, int j, uint8_t code) {  // 0-based, both directions
        sp[i].push_back({j, code});
        sp[j].push_back({i, code});
    };
    // short-range lists: for |j-i| <= max_nbr_range only listex/list14 are consulted (L3012-3031),
    // exclusion wins over 1-4
    for (int i = 0; i < ns; i++)
        for (int k = 1; k <= R && i + k < ns; k++) {
            size_t idx = (size_t)(k - 1) + (size_t)i * R;
            if (sys->listex[idx]) add(i, i + k, kPairExcluded);
            else if (sys->list14[idx]) add(i, i + k, kPair14);
        }
    // long-range lists only apply beyond max_nbr_range (L3032-3052); exclusion first
    std::vector<std::pair<int, int>> exl;
    for (int n = 0; n < s.nexlong; n++) {
        int a = sys->listexlong[2 * n] - 1, b = sys->listexlong[2 * n + 1] - 1;
        if (a < 0 || b < 0 || a >= ns || b >= ns || std::abs(a - b) <= R) continue;
        exl.push_back({std::min(a, b), std::max(a, b)});
    }
    std::sort(exl.begin(), exl.end());
    exl.erase(std::unique(exl.begin(), exl.end()), exl.end());
    for (auto &p : exl) add(p.first, p.second, kPairExcluded);
    std::vector<std::pair<int, int>> l14;
    for (int n = 0; n < s.n14long; n++) {
        int a = sys->list14long[2 * n] - 1, b = sys->list14long[2 * n + 1] - 1;
        if (a < 0 || b < 0 || a >= ns || b >= ns || std::abs(a - b) <= R) continue;
        std::pair<int, int> p{std::min(a, b), std::max(a, b)};
        if (std::binary_search(exl.begin(), exl.end(), p)) continue;
        l14.push_back(p);
    }
    std::sort(l14.begin(), l14.end());
    l14.erase(std::unique(l14.begin(), l14.end()), l14.end());
    for (auto &p : l14) add(p.first, p.second, kPair14);

    sp_off.assign(ns + 1, 0);
    for (int i = 0; i < ns; i++) {
        std::sort(sp[i].begin(), sp[i].end());
        sp_off[i + 1] = sp_off[i] + (int)sp[i].size();
    }
    sp_partner.resize(sp_off[ns]);
    sp_code.resize(sp_off[ns]);
    for (int i = 0; i < ns; i++)
        for (size_t k = 0; k < sp[i].size(); k++) {
            sp_partner[sp_off[i] + k] = sp[i][k].first;
            sp_code[sp_off[i] + k] = sp[i][k].second;
        }
    // per solute group: union of its atoms' special partners plus its own atoms
    gs_off.assign(s.ncgp_solute + 1, 0);
    gs_atoms.clear();
    for (int g = 0; g < s.ncgp_solute; g++) {
        std::vector<int> u;
        for (int k = 0; k < g_n[g]; k++) {
            int a = g_atoms[g_first[g] + k];
            u.push_back(a);
            if (a < ns)
                for (int m = sp_off[a]; m < sp_off[a + 1]; m++) u.push_back(sp_partner[m]);
        }
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        gs_atoms.insert(gs_atoms.end(), u.begin(), u.end());
        gs_off[g + 1] = (int)gs_atoms.size();
    }
    nq_off.assign(s.ncgp_solute + 1, 0);
    nq_atoms.clear();
    for (int g = 0; g < s.ncgp_solute; g++) {
        for (int k = 0; k < g_n[g]; k++)
            if (!is_q[g_atoms[g_first[g] + k]]) nq_atoms.push_back(g_atoms[g_first[g] + k]);
        nq_off[g + 1] = (int)nq_atoms.size();
    }
}

// precompute_set_values_qp (simprep.f90:3388-3440), arguments 1-based
inline QPar HostTables::qp_values(int iq, int j, int is, int vdw, bool &set) const {
    QPar p{};
    const int i = iqseq0[iq - 1] + 1;
    const int qvdw = (vdw == 2) ? 1 : vdw;
    const int tj = iac0[j - 1] + 1, ti = iac0[i - 1] + 1;
    double ai, bi;
    if (s.qvdw_flag) {
        ai = qavdw(qiac(iq, is), qvdw);
        bi = qbvdw(qiac(iq, is), qvdw);
    } else {
        ai = avdw(ti, vdw);
        bi = bvdw(ti, vdw);
    }
    combine(ai, bi, avdw(tj, vdw), bvdw(tj, vdw), p.A, p.B);
    p.el = (s.qq_use_library_charges ? crg[i - 1] : qcrg(iq, is)) * crg[j - 1];
    set = is_set(p);
    p.score = sc_lookup(iq, tj, is);
    if (vdw == 3) p.el *= s.el14_scale;
    return p;
}

inline void HostTables::build_q(const qnb_system *sys) {
    const int nq = s.nqat, nst = s.nstates, ns = s.nat_solute, sa = s.solv_atom;
    iqseq0.resize(nq);
    for (int q = 0; q < nq; q++) iqseq0[q] = sys->iqseq[q] - 1;
    qbonded.assign(ns, 0);
    for (int i = 1; i <= ns; i++) {
        bool any = false;
        for (int q = 1; q <= nq && !any; q++)
            for (int st = 1; st <= nst; st++)
                if (qconn(st, i, q) <= 3) { any = true; break; }
        qbonded[i - 1] = any;
    }
    if (nq == 0) return;
    qp_tab.assign((size_t)nq * nst * std::max(ns, 1), QPar{});
    qp_set.assign((size_t)nq * nst * std::max(ns, 1), 0);
    auto QP = [&](int iq, int is, int j) -> size_t { return ((size_t)(iq - 1) * nst + (is - 1)) * ns + (j - 1); };
    for (int j = 1; j <= ns; j++) {
        if (is_q[j - 1]) continue;
        if (!qbonded[j - 1]) {
            // qp_int_comp (simprep.f90:3131-3147).  The LJ code is chosen before the state loop and, once
            // switched to 3 by a 1-4 relation in one state, stays 3 for the later states.
            for (int q = 1; q <= nq; q++) {
                int vdw = ljcod(iac0[j - 1] + 1, iac0[iqseq0[q - 1]] + 1);
                for (int st = 1; st <= nst; st++) {
                    if (qconn(st, j, q) == 4) vdw = 3;
                    bool set;
                    qp_tab[QP(q, st, j)] = qp_values(q, j, st, vdw, set);
                    qp_set[QP(q, st, j)] = set;
                }
            }
        } else {
            // neighbours of Q-atoms: second half of qq_int_comp (simprep.f90:3232-3259)
            for (int q = 1; q <= nq; q++)
                for (int st = 1; st <= nst; st++) {
                    int c = qconn(st, j, q);
                    if (c < 4) continue;
                    int vdw = (c == 4) ? 3 : (s.qvdw_flag ? 1 : ljcod(iac0[iqseq0[q - 1]] + 1, iac0[j - 1] + 1));
                    bool set;
                    qp_tab[QP(q, st, j)] = qp_values(q, j, st, vdw, set);
                    qp_set[QP(q, st, j)] = set;
                }
        }
    }
    // static Q-Q list: qq_int_comp first half (simprep.f90:3172-3229) + nbqqlist (nonbondene.f90:3238-3254)
    qq_list.clear();
    qqp_list.clear();
    for (int q = 1; q <= nq - 1; q++)
        for (int r = q + 1; r <= nq; r++)
            for (int st = 1; st <= nst; st++) {
                const int ia = iqseq0[q - 1] + 1, ja = iqseq0[r - 1] + 1;
                const int c = qconn(st, ja, q);
                if (c < 4) continue;
                double elscale = 1.0;
                for (int e = 0; e < s.nel_scale; e++) {
                    int k = sys->el_scale_iq[e], l = sys->el_scale_jq[e];
                    if ((q == k && r == l) || (q == l && r == k)) {
                        elscale = sys->el_scale[e + (size_t)(st - 1) * s.nel_scale];
                        break;  // first match wins (do while .not.found)
                    }
                }
                int vdw;
                if (c == 4) vdw = 3;
                else if (!s.qvdw_flag) vdw = ljcod(iac0[ia - 1] + 1, iac0[ja - 1] + 1);
                else {
                    vdw = 1;
                    for (int e = 0; e < s.nqexpnb; e++)
                        if ((q == sys->iqexpnb[e] && r == sys->jqexpnb[e]) || (r == sys->iqexpnb[e] && q == sys->jqexpnb[e])) {
                            vdw = 2;
                            break;
                        }
                }
                // precompute_set_values_qq (simprep.f90:3444-3507)
                QPar p{};
                const bool softpair = (vdw == 2) && s.qvdw_flag;
                if (s.qvdw_flag) {
                    double ai = qavdw(qiac(q, st), vdw), aj = qavdw(qiac(r, st), vdw);
                    double bi = qbvdw(qiac(q, st), vdw), bj = qbvdw(qiac(r, st), vdw);
                    if (s.ivdw_rule == QNB_VDW_ARITHMETIC && softpair) {
                        p.A = ai * aj;   // soft pair keeps the plain products
                        p.B = bi * bj;
                    } else combine(ai, bi, aj, bj, p.A, p.B);
                } else {
                    combine(avdw(iac0[ia - 1] + 1, vdw), bvdw(iac0[ia - 1] + 1, vdw), avdw(iac0[ja - 1] + 1, vdw),
                            bvdw(iac0[ja - 1] + 1, vdw), p.A, p.B);
                }
                p.el = s.qq_use_library_charges ? crg[ia - 1] * crg[ja - 1] : qcrg(q, st) * qcrg(r, st);
                p.el *= elscale;
                p.score = sc_lookup(q, s.natyps + r, st);
                bool set = is_set(p);
                if (vdw == 3) p.el *= s.el14_scale;
                if (!set) continue;
                qq_list.push_back(StaticQPair{ia - 1, ja - 1, q, r, st - 1, softpair ? 1 : 0, p});
            }
    // static Q - neighbour list (nonbondene.f90:3257-3280)
    for (int j = 1; j <= ns; j++) {
        if (is_q[j - 1] || !qbonded[j - 1]) continue;
        for (int q = 1; q <= nq; q++)
            for (int st = 1; st <= nst; st++) {
                if (!qp_set[QP(q, st, j)]) continue;
                qqp_list.push_back(StaticQPair{iqseq0[q - 1], j - 1, q, 0, st - 1, 0, qp_tab[QP(q, st, j)]});
            }
    }
    // Q - water: qw_int_comp / precompute_set_values_qw (simprep.f90:3272-3292, 3511-3559)
    qw_tab.assign((size_t)nq * nst * std::max(sa, 1), QPar{});
    if (s.nwat > 0)
        for (int q = 1; q <= nq; q++)
            for (int site = 1; site <= sa; site++) {
                const int tw = iac0[ns + site - 1] + 1, ti = iac0[iqseq0[q - 1]] + 1;
                const int vdw = ljcod(tw, ti), qvdw = (vdw == 2) ? 1 : vdw;
                for (int st = 1; st <= nst; st++) {
                    QPar p{};
                    double ai, bi;
                    if (s.qvdw_flag) {
                        ai = qavdw(qiac(q, st), qvdw);
                        bi = qbvdw(qiac(q, st), qvdw);
                    } else {
                        ai = avdw(ti, vdw);
                        bi = bvdw(ti, vdw);
                    }
                    combine(ai, bi, avdw(tw, vdw), bvdw(tw, vdw), p.A, p.B);
                    p.el = (s.qq_use_library_charges ? crg[iqseq0[q - 1]] : qcrg(q, st)) * w_crg[site - 1];
                    p.score = sc_lookup(q, tw, st);
                    qw_tab[((size_t)(q - 1) * nst + (st - 1)) * sa + (site - 1)] = p;
                }
            }
}

inline bool HostTables::build(const qnb_system *sys) {
    s = *sys;
    s_iaclib = sys->iaclib; s_qcrg = sys->qcrg; s_qavdw = sys->qavdw; s_qbvdw = sys->qbvdw; s_sc = sys->sc_lookup;
    s_ljcod = sys->ljcod; s_qconn = sys->qconn; s_qiac = sys->qiac;
    const int n = s.natom, ns = s.nat_solute;
    if (s.nstates < 1 || s.nstates > kMaxStates) { error = "nstates must be 1.." + std::to_string(kMaxStates); return false; }
    if (s.nwat > 0 && s.solvent_type == QNB_SOLVENT_GENERAL) {
        error = "Topology contains mixed solvent. This feature is not implemented yet.";  // simprep.f90:3620
        return false;
    }
    if (s.nwat > 0 && s.solv_atom != 3) { error = "only 3-site solvent is supported by the CUDA water kernels"; return false; }
    if (s.ntors_gt_solute) { error = "solvents with internal torsions (nonbond_solvent_internal) are not supported"; return false; }
    if (n >= (1 << 24)) { error = "more than 2^24 atoms: row entries hold 24-bit packed atom indices"; return false; }
    if (s.use_PBC && s.nqat > 0 && (s.qswitch < 1 || s.qswitch > n)) { error = "PBC with Q-atoms needs a valid qswitch"; return false; }

    iac0.resize(n);
    crg.assign(sys->crg, sys->crg + n);
    is_q.resize(n);
    excl.resize(n);
    for (int i = 0; i < n; i++) {
        iac0[i] = sys->iac[i] - 1;
        is_q[i] = sys->iqatom[i] != 0;
        excl[i] = (!s.use_PBC) && sys->excl[i] != 0;   // topo_read clears excl for PBC (topo.f90:1053)
    }
    // compact atom types
    std::vector<int> map(s.natyps > s.num_atyp ? s.natyps : s.num_atyp, -1);
    ctype.resize(n);
    ct_to_iac.clear();
    for (int i = 0; i < n; i++) {
        int t = iac0[i];
        if (t < 0 || t >= (int)map.size() || t >= s.natyps) { error = "atom type code out of range"; return false; }
        if (map[t] < 0) { map[t] = (int)ct_to_iac.size(); ct_to_iac.push_back(t + 1); }
        ctype[i] = map[t];
    }
    nct = (int)ct_to_iac.size();
    lj_a.resize(nct * 3);
    lj_b.resize(nct * 3);
    ljcode.resize((size_t)nct * nct);
    for (int a = 0; a < nct; a++) {
        for (int c = 1; c <= 3; c++) {
            lj_a[a * 3 + c - 1] = avdw(ct_to_iac[a], c);
            lj_b[a * 3 + c - 1] = bvdw(ct_to_iac[a], c);
        }
        for (int b = 0; b < nct; b++) ljcode[(size_t)a * nct + b] = (uint8_t)ljcod(ct_to_iac[a], ct_to_iac[b]);
    }
    // charge groups
    g_first.resize(s.ncgp); g_n.resize(s.ncgp); g_switch.resize(s.ncgp);
    g_atoms.resize(n);
    grp_of_atom.assign(n, -1);
    for (int i = 0; i < n; i++) g_atoms[i] = sys->cgpatom[i] - 1;
    for (int g = 0; g < s.ncgp; g++) {
        g_switch[g] = sys->cgp[3 * g] - 1;
        g_first[g] = sys->cgp[3 * g + 1] - 1;
        g_n[g] = sys->cgp[3 * g + 2] - sys->cgp[3 * g + 1] + 1;
        for (int k = 0; k < g_n[g]; k++) grp_of_atom[g_atoms[g_first[g] + k]] = g;   // iwhich_cgp, simprep.f90:3660
    }
    // units
    nunit = s.ncgp_solute + s.nwat;
    u_sw.resize(nunit); u_grp.resize(nunit); u_excl.resize(nunit);
    std::vector<int> seen(s.ncgp, 0);
    for (int u = 0; u < nunit; u++) {
        if (u < s.ncgp_solute) { u_sw[u] = g_switch[u]; u_grp[u] = u; }
        else {
            int w = u - s.ncgp_solute;
            u_sw[u] = ns + s.solv_atom * w;   // first atom of the molecule (nonbondene.f90:4121)
            u_grp[u] = grp_of_atom[u_sw[u]];
            if (u_grp[u] < 0) { error = "water molecule without charge group"; return false; }
        }
        if (seen[u_grp[u]]++) { error = "two units share one charge group"; return false; }
        if (g_switch[u_grp[u]] != u_sw[u]) { error = "a solvent charge group whose switching atom is not the molecule's first atom is not supported"; return false; }
        u_excl[u] = excl[u_sw[u]];
    }
    // water sites (simprep.f90:3610-3621)
    w_crg.assign(s.solv_atom, 0.0);
    w_ctype.assign(s.solv_atom, 0);
    ww_par.assign((size_t)s.solv_atom * s.solv_atom, QPar{});
    if (s.nwat > 0) {
        for (int k = 0; k < s.solv_atom; k++) { w_crg[k] = crg[ns + k]; w_ctype[k] = ctype[ns + k]; }
        for (int a = 0; a < s.solv_atom; a++)
            for (int b = 0; b < s.solv_atom; b++) {   // precompute_set_values_ww (simprep.f90:3563-3589)
                QPar p{};
                int code = ljcode[(size_t)w_ctype[a] * nct + w_ctype[b]];
                combine(lj_a[w_ctype[a] * 3 + code - 1], lj_b[w_ctype[a] * 3 + code - 1],
                        lj_a[w_ctype[b] * 3 + code - 1], lj_b[w_ctype[b] * 3 + code - 1], p.A, p.B);
                p.el = w_crg[a] * w_crg[b];
                ww_par[(size_t)a * s.solv_atom + b] = p;
            }
    }
    build_specials(sys);
    build_q(sys);
    s_iaclib = s_qcrg = s_qavdw = s_qbvdw = s_sc = nullptr;
    s_ljcod = s_qconn = s_qiac = nullptr;
    return true;
}

}  // namespace qnb
