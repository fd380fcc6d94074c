/*
 * qoracle.c -- TEST INFRASTRUCTURE ONLY (see qoracle.h).
 *
 * Literal FP64 restatement of the Qdyn6 nonbonded path.  Reference = qusers/Q6
 * src/ (.f90); every function names the lines it follows.  Only the serial
 * (#else, non-OpenMP) branches are restated: the OpenMP branches are not built
 * by any default target and are buggy.
 *
 * Indices: the qnb_system tables carry 1-based indices (as the Fortran host has
 * them); this file keeps them 1-based in loops so the code reads like the
 * reference, and subtracts 1 only when touching C arrays.
 */
#define _POSIX_C_SOURCE 200809L
#include "qoracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define QREAL_EPS 1e-10 /* sizes.f90:50 (QDOUBLE) */

typedef struct { double x, y, z; } vec3;

/* globals.f90:320-331 */
typedef struct { double vdWA, vdWB, elec, score; int set; } precomp_t;
typedef struct { double vdWA, vdWB, elec, score; int set, soft; } precomp_qq_t;
/* globals.f90:367-389 */
typedef struct { int i, j, cgp_pair; double vdWA, vdWB, elec; } nb_t;
typedef struct { int i, j; vec3 shift; } cgp_pair_t;
typedef struct { int i, j, cgp_pair; double vdWA, vdWB, elec, score; } nbqp_t;
typedef struct { int iq, jq, soft; double vdWA, vdWB, elec, score; } nbq_t;
/* globals.f90:503-509 */
typedef struct { vec3 cgp_cent; double phi0; vec3 phi1; vec3 phi2[3]; vec3 phi3[9]; } lrf_t;
/* globals.f90:392-395 */
typedef struct { double Vel, V_a, V_b, dv; vec3 vec; } eneret_t;

typedef struct { nb_t *p; int64_t n, cap; } nb_list;
typedef struct { cgp_pair_t *p; int64_t n, cap; } cgp_list;

struct qo_state {
    qnb_system s; /* scalars + pointers into owned copies below */
    /* owned copies */
    int32_t *cgp, *cgpatom, *excl, *iqatom, *iqseq, *iac, *ljcod, *listex, *list14, *listexlong, *list14long;
    int32_t *qiac, *iqexpnb, *jqexpnb, *el_scale_iq, *el_scale_jq, *qconn;
    double *crg, *iaclib, *qcrg, *qavdw, *qbvdw, *sc_lookup, *el_scale;
    /* derived */
    int32_t *iwhich_cgp;       /* simprep.f90:3660-3667 */
    unsigned char *qbonded;    /* any(qconn(:,i,:) <= 3) per solute atom */
    double *chg_solv, *aLJ_solv, *bLJ_solv; /* simprep.f90:3610-3621 */
    vec3 boxlength, inv_boxl;
    /* precompute tables */
    precomp_t *pw_precomp;  /* [nat_solute][solv_atom] */
    precomp_t *ww_precomp;  /* [solv_atom][solv_atom] */
    precomp_t *qp_precomp;  /* [nat_solute][nqat][nstates] */
    precomp_t *qw_precomp;  /* [nqat][solv_atom][nstates] */
    precomp_qq_t *qq_precomp; /* [nqat][nqat][nstates] */
    /* lists */
    nb_list nbpp, nbpw, nbww;
    cgp_list nbpp_cgp, nbpw_cgp, nbqp_cgp;
    nbqp_t *nbqp; int64_t nbqp_pair, nbqp_cap; /* [pair][state] */
    nbqp_t *nbqw; int64_t nbqw_pair, nbqw_cap;
    nbq_t *nbqq; int *nbqq_pair; int nbqq_max;   /* [state][pair] */
    nbqp_t *nbqqp; int *nbqqp_pair;
    int qp_list_done, qw_list_done;
    lrf_t *lrf;
    double RcLRF_global;
};

static __thread char qo_err[512];
const char *qo_last_error(void) { return qo_err; }

/* ------------------------------------------------------------------ helpers */
static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "qoracle: out of memory\n"); abort(); }
    return p;
}
static void *dup_arr(const void *src, size_t n, size_t sz) {
    void *p = xcalloc(n, sz);
    if (src && n) memcpy(p, src, n * sz);
    return p;
}
static inline vec3 v_sub(vec3 a, vec3 b) { vec3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline vec3 v_add(vec3 a, vec3 b) { vec3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline vec3 v_scale(vec3 a, double f) { vec3 r = {a.x * f, a.y * f, a.z * f}; return r; }
static inline vec3 v_mul(vec3 a, vec3 b) { vec3 r = {a.x * b.x, a.y * b.y, a.z * b.z}; return r; }
static inline double v_dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* math.f90:206-212 qvec_square */
static inline double qvec_square(vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
/* math.f90:254-262 q_dist4 */
static inline double q_dist4(vec3 a, vec3 b) { return qvec_square(v_sub(b, a)); }
/* math.f90:566-574 q_nint: Fortran nint rounds half away from zero == C round() */
static inline vec3 q_nint(vec3 a) { vec3 r = {round(a.x), round(a.y), round(a.z)}; return r; }
static inline vec3 X(const double *x, int i /*1-based*/) {
    vec3 r = {x[3 * (i - 1)], x[3 * (i - 1) + 1], x[3 * (i - 1) + 2]};
    return r;
}
static inline void d_sub(double *d, int i, vec3 v) { d[3*(i-1)] -= v.x; d[3*(i-1)+1] -= v.y; d[3*(i-1)+2] -= v.z; }
static inline void d_add(double *d, int i, vec3 v) { d[3*(i-1)] += v.x; d[3*(i-1)+1] += v.y; d[3*(i-1)+2] += v.z; }

#define S (st->s)
#define CGP_ISWITCH(ig) (st->cgp[3 * ((ig)-1) + 0])
#define CGP_FIRST(ig) (st->cgp[3 * ((ig)-1) + 1])
#define CGP_LAST(ig) (st->cgp[3 * ((ig)-1) + 2])
#define CGPATOM(ia) (st->cgpatom[(ia)-1])
#define EXCL(i) (st->excl[(i)-1] != 0)
#define IQATOM(i) (st->iqatom[(i)-1])
#define IQSEQ(iq) (st->iqseq[(iq)-1])
#define IAC(i) (st->iac[(i)-1])
#define CRG(i) (st->crg[(i)-1])
#define AVDW(t, c) (st->iaclib[7 * ((t)-1) + 1 + ((c)-1)])
#define BVDW(t, c) (st->iaclib[7 * ((t)-1) + 4 + ((c)-1)])
#define LJCOD(a, b) (st->ljcod[((a)-1) + (size_t)((b)-1) * S.num_atyp])
#define LISTEX(k, i) (st->listex[((k)-1) + (size_t)((i)-1) * S.max_nbr_range] != 0)
#define LIST14(k, i) (st->list14[((k)-1) + (size_t)((i)-1) * S.max_nbr_range] != 0)
#define QCRG(iq, is) (st->qcrg[((iq)-1) + (size_t)((is)-1) * S.nqat])
#define QIAC(iq, is) (st->qiac[((iq)-1) + (size_t)((is)-1) * S.nqat])
#define QAVDW(t, c) (st->qavdw[((t)-1) + (size_t)((c)-1) * S.nqlib])
#define QBVDW(t, c) (st->qbvdw[((t)-1) + (size_t)((c)-1) * S.nqlib])
#define SC_LOOKUP(iq, k, is) (st->sc_lookup[((iq)-1) + (size_t)((k)-1) * S.nqat + (size_t)((is)-1) * S.nqat * (S.natyps + S.nqat)])
#define QCONN(is, i, iq) (st->qconn[((is)-1) + (size_t)((i)-1) * S.nstates + (size_t)((iq)-1) * S.nstates * S.nat_solute])
#define CHG_SOLV(j) (st->chg_solv[(j)-1])
#define ALJ_SOLV(j, c) (st->aLJ_solv[((j)-1) + ((c)-1) * S.solv_atom])
#define BLJ_SOLV(j, c) (st->bLJ_solv[((j)-1) + ((c)-1) * S.solv_atom])
#define PW_PRECOMP(i, j) (st->pw_precomp[(size_t)((i)-1) * S.solv_atom + ((j)-1)])
#define WW_PRECOMP(a, b) (st->ww_precomp[((a)-1) * S.solv_atom + ((b)-1)])
#define QP_PRECOMP(j, iq, is) (st->qp_precomp[((size_t)((j)-1) * S.nqat + ((iq)-1)) * S.nstates + ((is)-1)])
#define QW_PRECOMP(iq, j, is) (st->qw_precomp[((size_t)((iq)-1) * S.solv_atom + ((j)-1)) * S.nstates + ((is)-1)])
#define QQ_PRECOMP(iq, jq, is) (st->qq_precomp[((size_t)((iq)-1) * S.nqat + ((jq)-1)) * S.nstates + ((is)-1)])
#define NBQP(ip, is) (st->nbqp[(size_t)((ip)-1) * S.nstates + ((is)-1)])
#define NBQW(ip, is) (st->nbqw[(size_t)((ip)-1) * S.nstates + ((is)-1)])
#define NBQQ(ip, is) (st->nbqq[(size_t)((is)-1) * st->nbqq_max + ((ip)-1)])
#define NBQQP(ip, is) (st->nbqqp[(size_t)((is)-1) * st->nbqq_max + ((ip)-1)])

static int is_set(double a, double b, double e) {
    /* simprep.f90:3345 */
    return (fabs(a) > QREAL_EPS) || (fabs(b) > QREAL_EPS) || (fabs(e) > QREAL_EPS);
}

/* ---------------------------------------------------------------- precompute */

/* precompute_set_values_pp, simprep.f90:3321-3351 */
static precomp_t precompute_set_values_pp(const qo_state *st, int i, int j, int vdw) {
    precomp_t p = {0, 0, 0, 0, 0};
    double tempA, tempB;
    if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
        tempA = AVDW(IAC(i), vdw) * AVDW(IAC(j), vdw);
        tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        p.vdWA = tempA;
        p.vdWB = tempB;
    } else {
        tempA = AVDW(IAC(i), vdw) + AVDW(IAC(j), vdw);
        tempA = tempA * tempA;
        tempA = tempA * tempA * tempA;
        tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        p.vdWA = (tempA * tempA) * tempB;
        p.vdWB = 2.0 * tempA * tempB;
    }
    p.elec = CRG(i) * CRG(j);
    if (is_set(p.vdWA, p.vdWB, p.elec)) p.set = 1;
    if (vdw == 3) p.elec = p.elec * S.el14_scale;
    return p;
}

/*
 * pp_int_comp, simprep.f90:2952-3065: what pp_map(i-pp_low,j)/pp_precomp hold for the
 * atom pair (i,j).  Returns 0 when pp_map would be 0 (excluded pair).  The map
 * itself (nat_solute^2 int32) is not materialised; the per-pair decision is the same.
 */
static int pp_lookup(const qo_state *st, int i, int j, precomp_t *out) {
    int nl;
    if (abs(j - i) <= S.max_nbr_range) {
        if (i < j) {
            if (LISTEX(j - i, i)) return 0;
            if (LIST14(j - i, i)) { *out = precompute_set_values_pp(st, i, j, 3); return 1; }
        } else {
            if (LISTEX(i - j, j)) return 0;
            if (LIST14(i - j, j)) { *out = precompute_set_values_pp(st, i, j, 3); return 1; }
        }
    } else {
        for (nl = 0; nl < S.nexlong; nl++) {
            int a = st->listexlong[2 * nl], b = st->listexlong[2 * nl + 1];
            if ((a == i && b == j) || (a == j && b == i)) return 0;
        }
        for (nl = 0; nl < S.n14long; nl++) {
            int a = st->list14long[2 * nl], b = st->list14long[2 * nl + 1];
            if ((a == i && b == j) || (a == j && b == i)) { *out = precompute_set_values_pp(st, i, j, 3); return 1; }
        }
    }
    *out = precompute_set_values_pp(st, i, j, LJCOD(IAC(i), IAC(j)));
    return 1;
}

/* precompute_set_values_pw, simprep.f90:3355-3384; pw_int_comp, simprep.f90:3069-3108 */
static void pw_int_comp(qo_state *st) {
    int i, j;
    st->pw_precomp = xcalloc((size_t)S.nat_solute * S.solv_atom, sizeof(precomp_t));
    for (i = 1; i <= S.nat_solute; i++) {
        if (IQATOM(i) != 0) continue;
        for (j = 1; j <= S.solv_atom; j++) {
            int vdw = LJCOD(IAC(i), IAC(S.nat_solute + j));
            precomp_t *p = &PW_PRECOMP(i, j);
            double tempA, tempB;
            if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
                tempA = AVDW(IAC(i), vdw) * ALJ_SOLV(j, vdw);
                tempB = BVDW(IAC(i), vdw) * BLJ_SOLV(j, vdw);
                p->vdWA = tempA;
                p->vdWB = tempB;
            } else {
                tempA = AVDW(IAC(i), vdw) + ALJ_SOLV(j, vdw);
                tempA = tempA * tempA;
                tempA = tempA * tempA * tempA;
                tempB = BVDW(IAC(i), vdw) * BLJ_SOLV(j, vdw);
                p->vdWA = (tempA * tempA) * tempB;
                p->vdWB = 2.0 * tempA * tempB;
            }
            p->elec = CRG(i) * CHG_SOLV(j);
            if (is_set(p->vdWA, p->vdWB, p->elec)) p->set = 1;
        }
    }
}

/* precompute_set_values_qp, simprep.f90:3388-3440 */
static void precompute_set_values_qp(qo_state *st, int iq, int j, int istate, int vdw) {
    int i = IQSEQ(iq), qvdw;
    double tempA, tempB;
    precomp_t *p = &QP_PRECOMP(j, iq, istate);
    qvdw = (vdw == 2) ? 1 : vdw;
    if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
        if (S.qvdw_flag) {
            tempA = QAVDW(QIAC(iq, istate), qvdw) * AVDW(IAC(j), vdw);
            tempB = QBVDW(QIAC(iq, istate), qvdw) * BVDW(IAC(j), vdw);
        } else {
            tempA = AVDW(IAC(i), vdw) * AVDW(IAC(j), vdw);
            tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        }
        p->vdWA = tempA;
        p->vdWB = tempB;
    } else {
        if (S.qvdw_flag) {
            tempA = QAVDW(QIAC(iq, istate), qvdw) + AVDW(IAC(j), vdw);
            tempB = QBVDW(QIAC(iq, istate), qvdw) * BVDW(IAC(j), vdw);
        } else {
            tempA = AVDW(IAC(i), vdw) + AVDW(IAC(j), vdw);
            tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        }
        tempA = tempA * tempA;
        tempA = tempA * tempA * tempA;
        p->vdWA = (tempA * tempA) * tempB;
        p->vdWB = 2.0 * tempA * tempB;
    }
    if (!S.qq_use_library_charges) p->elec = QCRG(iq, istate) * CRG(j);
    else p->elec = CRG(i) * CRG(j);
    if (is_set(p->vdWA, p->vdWB, p->elec)) p->set = 1;
    p->score = SC_LOOKUP(iq, IAC(j), istate);
    if (vdw == 3) p->elec = p->elec * S.el14_scale;
}

/* qp_int_comp, simprep.f90:3112-3148 */
static void qp_int_comp(qo_state *st) {
    int ig, jq, is, vdw;
    st->qp_precomp = xcalloc((size_t)S.nat_solute * S.nqat * S.nstates, sizeof(precomp_t));
    for (ig = 1; ig <= S.nat_solute; ig++) {
        if (st->qbonded[ig - 1]) continue;
        for (jq = 1; jq <= S.nqat; jq++) {
            vdw = LJCOD(IAC(ig), IAC(IQSEQ(jq)));
            for (is = 1; is <= S.nstates; is++) {
                /* NB: vdw stays 3 for later states once set (reference behaviour, L3134-3137) */
                if (QCONN(is, ig, jq) == 4) vdw = 3;
                precompute_set_values_qp(st, jq, ig, is, vdw);
            }
        }
    }
}

/* precompute_set_values_qq, simprep.f90:3444-3507 */
static void precompute_set_values_qq(qo_state *st, int iq, int jq, int istate, int vdw, double q_elscale) {
    int i = IQSEQ(iq), j = IQSEQ(jq);
    double tempA, tempB;
    precomp_qq_t *p = &QQ_PRECOMP(iq, jq, istate);
    if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
        if (S.qvdw_flag) {
            tempA = QAVDW(QIAC(iq, istate), vdw) * QAVDW(QIAC(jq, istate), vdw);
            tempB = QBVDW(QIAC(iq, istate), vdw) * QBVDW(QIAC(jq, istate), vdw);
        } else {
            tempA = AVDW(IAC(i), vdw) * AVDW(IAC(j), vdw);
            tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        }
        p->vdWA = tempA;
        p->vdWB = tempB;
    } else {
        if (S.qvdw_flag) {
            if (vdw == 2) tempA = QAVDW(QIAC(iq, istate), vdw) * QAVDW(QIAC(jq, istate), vdw);
            else tempA = QAVDW(QIAC(iq, istate), vdw) + QAVDW(QIAC(jq, istate), vdw);
            tempB = QBVDW(QIAC(iq, istate), vdw) * QBVDW(QIAC(jq, istate), vdw);
        } else {
            tempA = AVDW(IAC(i), vdw) + AVDW(IAC(j), vdw);
            tempB = BVDW(IAC(i), vdw) * BVDW(IAC(j), vdw);
        }
        if (vdw == 2 && S.qvdw_flag) {
            p->vdWA = tempA;
            p->vdWB = tempB;
        } else {
            tempA = tempA * tempA;
            tempA = tempA * tempA * tempA;
            p->vdWA = (tempA * tempA) * tempB;
            p->vdWB = 2.0 * tempA * tempB;
        }
    }
    if (!S.qq_use_library_charges) p->elec = QCRG(iq, istate) * QCRG(jq, istate);
    else p->elec = CRG(i) * CRG(j);
    p->elec = p->elec * q_elscale;
    p->score = SC_LOOKUP(iq, S.natyps + jq, istate);
    if (is_set(p->vdWA, p->vdWB, p->elec)) p->set = 1;
    if (vdw == 3) p->elec = p->elec * S.el14_scale;
    if (vdw == 2 && S.qvdw_flag) p->soft = 1;
}

/* qq_int_comp, simprep.f90:3152-3268 (without the excluded-group bookkeeping) */
static void qq_int_comp(qo_state *st) {
    int iq, jq, ia, ja, is, vdw, i, k, l, found;
    double tmp_elscale;
    st->qq_precomp = xcalloc((size_t)S.nqat * S.nqat * S.nstates, sizeof(precomp_qq_t));
    for (iq = 1; iq <= S.nqat - 1; iq++) {
        ia = IQSEQ(iq);
        for (jq = iq + 1; jq <= S.nqat; jq++) {
            ja = IQSEQ(jq);
            for (is = 1; is <= S.nstates; is++) {
                if (QCONN(is, ja, iq) >= 4) {
                    tmp_elscale = 1.0;
                    if (S.nel_scale != 0) {
                        i = 1; found = 0;
                        while (i <= S.nel_scale && !found) {
                            k = st->el_scale_iq[i - 1];
                            l = st->el_scale_jq[i - 1];
                            if ((iq == k && jq == l) || (iq == l && jq == k)) {
                                tmp_elscale = st->el_scale[(i - 1) + (size_t)(is - 1) * S.nel_scale];
                                found = 1;
                            }
                            i++;
                        }
                    }
                    if (QCONN(is, ja, iq) == 4) vdw = 3;
                    else if (!S.qvdw_flag) vdw = LJCOD(IAC(ia), IAC(ja));
                    else {
                        vdw = 1; i = 1; found = 0;
                        while (i <= S.nqexpnb && !found) {
                            if ((iq == st->iqexpnb[i - 1] && jq == st->jqexpnb[i - 1]) ||
                                (jq == st->iqexpnb[i - 1] && iq == st->jqexpnb[i - 1])) { vdw = 2; found = 1; }
                            i++;
                        }
                    }
                    precompute_set_values_qq(st, iq, jq, is, vdw, tmp_elscale);
                }
            }
        }
    }
    /* Q - bonded-neighbour non-Q atoms go to qp_precomp, L3232-3259 */
    for (ja = 1; ja <= S.nat_solute; ja++) {
        if (IQATOM(ja) != 0) continue;
        if (!st->qbonded[ja - 1]) continue;
        for (iq = 1; iq <= S.nqat; iq++) {
            ia = IQSEQ(iq);
            for (is = 1; is <= S.nstates; is++) {
                if (QCONN(is, ja, iq) >= 4) {
                    if (QCONN(is, ja, iq) == 4) vdw = 3;
                    else if (S.qvdw_flag) vdw = 1;
                    else vdw = LJCOD(IAC(ia), IAC(ja));
                    precompute_set_values_qp(st, iq, ja, is, vdw);
                }
            }
        }
    }
}

/* precompute_set_values_qw, simprep.f90:3511-3559; qw_int_comp, simprep.f90:3272-3292 */
static void qw_int_comp(qo_state *st) {
    int iq, j, istate;
    st->qw_precomp = xcalloc((size_t)S.nqat * S.solv_atom * S.nstates, sizeof(precomp_t));
    for (iq = 1; iq <= S.nqat; iq++) {
        for (j = 1; j <= S.solv_atom; j++) {
            int vdw = LJCOD(IAC(S.nat_solute + j), IAC(IQSEQ(iq)));
            int qvdw = (vdw == 2) ? 1 : vdw;
            int i = IQSEQ(iq), iacj = IAC(S.nat_solute + j);
            for (istate = 1; istate <= S.nstates; istate++) {
                precomp_t *p = &QW_PRECOMP(iq, j, istate);
                double tempA, tempB;
                if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
                    if (S.qvdw_flag) {
                        tempA = QAVDW(QIAC(iq, istate), qvdw) * ALJ_SOLV(j, vdw);
                        tempB = QBVDW(QIAC(iq, istate), qvdw) * BLJ_SOLV(j, vdw);
                    } else {
                        tempA = AVDW(IAC(i), vdw) * ALJ_SOLV(j, vdw);
                        tempB = BVDW(IAC(i), vdw) * BLJ_SOLV(j, vdw);
                    }
                    p->vdWA = tempA;
                    p->vdWB = tempB;
                } else {
                    if (S.qvdw_flag) {
                        tempA = QAVDW(QIAC(iq, istate), qvdw) + ALJ_SOLV(j, vdw);
                        tempB = QBVDW(QIAC(iq, istate), qvdw) * BLJ_SOLV(j, vdw);
                    } else {
                        tempA = AVDW(IAC(i), vdw) + ALJ_SOLV(j, vdw);
                        tempB = BVDW(IAC(i), vdw) * BLJ_SOLV(j, vdw);
                    }
                    tempA = tempA * tempA;
                    tempA = tempA * tempA * tempA;
                    p->vdWA = (tempA * tempA) * tempB;
                    p->vdWB = 2.0 * tempA * tempB;
                }
                if (!S.qq_use_library_charges) p->elec = QCRG(iq, istate) * CHG_SOLV(j);
                else p->elec = CRG(i) * CHG_SOLV(j);
                p->score = SC_LOOKUP(iq, iacj, istate);
            }
        }
    }
}

/* precompute_set_values_ww, simprep.f90:3563-3589; ww_int_comp, simprep.f90:3296-3317 */
static void ww_int_comp(qo_state *st) {
    int i, j;
    st->ww_precomp = xcalloc((size_t)S.solv_atom * S.solv_atom, sizeof(precomp_t));
    for (i = 1; i <= S.solv_atom; i++)
        for (j = 1; j <= S.solv_atom; j++) {
            int vdw = LJCOD(IAC(S.nat_solute + i), IAC(S.nat_solute + j));
            precomp_t *p = &WW_PRECOMP(i, j);
            double tempA, tempB;
            if (S.ivdw_rule == QNB_VDW_GEOMETRIC) {
                tempA = ALJ_SOLV(i, vdw) * ALJ_SOLV(j, vdw);
                tempB = BLJ_SOLV(i, vdw) * BLJ_SOLV(j, vdw);
                p->vdWA = tempA;
                p->vdWB = tempB;
            } else {
                tempA = ALJ_SOLV(i, vdw) + ALJ_SOLV(j, vdw);
                tempA = tempA * tempA;
                tempA = tempA * tempA * tempA;
                tempB = BLJ_SOLV(i, vdw) * BLJ_SOLV(j, vdw);
                p->vdWA = (tempA * tempA) * tempB;
                p->vdWB = 2.0 * tempA * tempB;
            }
            p->set = 1;
            p->elec = CHG_SOLV(i) * CHG_SOLV(j);
        }
}

/* nbqq_count, nonbondene.f90:3193-3227 */
static int nbqq_count(qo_state *st) {
    int iq, jq, is, j, m = 0;
    int *cnt = xcalloc(S.nstates, sizeof(int));
    for (iq = 1; iq <= S.nqat - 1; iq++)
        for (jq = iq + 1; jq <= S.nqat; jq++)
            for (is = 1; is <= S.nstates; is++)
                if (QCONN(is, IQSEQ(jq), iq) > 3) cnt[is - 1]++;
    for (j = 1; j <= S.nat_solute; j++) {
        if (IQATOM(j) > 0) continue;
        if (st->qbonded[j - 1])
            for (iq = 1; iq <= S.nqat; iq++)
                for (is = 1; is <= S.nstates; is++)
                    if (QCONN(is, j, iq) >= 4) cnt[is - 1]++;
    }
    for (is = 0; is < S.nstates; is++) if (cnt[is] > m) m = cnt[is];
    free(cnt);
    return m;
}

/* nbqqlist, nonbondene.f90:3231-3282 */
static void nbqqlist(qo_state *st) {
    int iq, jq, is, ja;
    for (is = 0; is < S.nstates; is++) { st->nbqq_pair[is] = 0; st->nbqqp_pair[is] = 0; }
    for (iq = 1; iq <= S.nqat - 1; iq++)
        for (jq = iq + 1; jq <= S.nqat; jq++)
            for (is = 1; is <= S.nstates; is++) {
                precomp_qq_t *q = &QQ_PRECOMP(iq, jq, is);
                nbq_t *e;
                if (!q->set) continue;
                st->nbqq_pair[is - 1]++;
                e = &NBQQ(st->nbqq_pair[is - 1], is);
                e->iq = iq; e->jq = jq; e->vdWA = q->vdWA; e->vdWB = q->vdWB; e->elec = q->elec;
                e->score = q->score; e->soft = q->soft;
            }
    for (ja = 1; ja <= S.nat_solute; ja++) {
        if (IQATOM(ja) != 0) continue;
        if (!st->qbonded[ja - 1]) continue;
        for (iq = 1; iq <= S.nqat; iq++)
            for (is = 1; is <= S.nstates; is++) {
                precomp_t *q = &QP_PRECOMP(ja, iq, is);
                nbqp_t *e;
                if (!q->set) continue;
                st->nbqqp_pair[is - 1]++;
                e = &NBQQP(st->nbqqp_pair[is - 1], is);
                e->i = iq; e->j = ja; e->vdWA = q->vdWA; e->vdWB = q->vdWB; e->elec = q->elec;
                e->score = q->score; e->cgp_pair = 0;
            }
    }
}

/* --------------------------------------------------------------- life cycle */
qo_state *qo_create(const qnb_system *sys) {
    qo_state *st;
    int i, ig, j, c;
    if (sys->abi_version != QNB_ABI_VERSION) { snprintf(qo_err, sizeof qo_err, "abi version mismatch"); return NULL; }
    if (sys->nwat > 0 && sys->solvent_type == QNB_SOLVENT_GENERAL) {
        /* simprep.f90:3620: die('Topology contains mixed solvent...') */
        snprintf(qo_err, sizeof qo_err, "Topology contains mixed solvent. This feature is not implemented yet.");
        return NULL;
    }
    if (sys->ntors_gt_solute) { snprintf(qo_err, sizeof qo_err, "nonbond_solvent_internal not restated"); return NULL; }
    st = xcalloc(1, sizeof *st);
    st->s = *sys;
    st->cgp = dup_arr(sys->cgp, 3 * (size_t)sys->ncgp, 4);
    st->cgpatom = dup_arr(sys->cgpatom, sys->natom, 4);
    st->excl = dup_arr(sys->excl, sys->natom, 4);
    st->iqatom = dup_arr(sys->iqatom, sys->natom, 4);
    st->iqseq = dup_arr(sys->iqseq, sys->nqat, 4);
    st->iac = dup_arr(sys->iac, sys->natom, 4);
    st->crg = dup_arr(sys->crg, sys->natom, 8);
    st->iaclib = dup_arr(sys->iaclib, 7 * (size_t)sys->natyps, 8);
    st->ljcod = dup_arr(sys->ljcod, (size_t)sys->num_atyp * sys->num_atyp, 4);
    st->listex = dup_arr(sys->listex, (size_t)sys->max_nbr_range * sys->nat_solute, 4);
    st->list14 = dup_arr(sys->list14, (size_t)sys->max_nbr_range * sys->nat_solute, 4);
    st->listexlong = dup_arr(sys->listexlong, 2 * (size_t)sys->nexlong, 4);
    st->list14long = dup_arr(sys->list14long, 2 * (size_t)sys->n14long, 4);
    st->qcrg = dup_arr(sys->qcrg, (size_t)sys->nqat * sys->nstates, 8);
    st->qiac = dup_arr(sys->qiac, (size_t)sys->nqat * sys->nstates, 4);
    st->qavdw = dup_arr(sys->qavdw, 3 * (size_t)sys->nqlib, 8);
    st->qbvdw = dup_arr(sys->qbvdw, 3 * (size_t)sys->nqlib, 8);
    st->sc_lookup = dup_arr(sys->sc_lookup, (size_t)sys->nqat * (sys->natyps + sys->nqat) * sys->nstates, 8);
    st->iqexpnb = dup_arr(sys->iqexpnb, sys->nqexpnb, 4);
    st->jqexpnb = dup_arr(sys->jqexpnb, sys->nqexpnb, 4);
    st->el_scale_iq = dup_arr(sys->el_scale_iq, sys->nel_scale, 4);
    st->el_scale_jq = dup_arr(sys->el_scale_jq, sys->nel_scale, 4);
    st->el_scale = dup_arr(sys->el_scale, (size_t)sys->nel_scale * sys->nstates, 8);
    st->qconn = dup_arr(sys->qconn, (size_t)sys->nstates * sys->nat_solute * sys->nqat, 4);

    /* iwhich_cgp, simprep.f90:3660-3667 */
    st->iwhich_cgp = xcalloc(S.natom, 4);
    for (ig = 1; ig <= S.ncgp; ig++)
        for (i = CGP_FIRST(ig); i <= CGP_LAST(ig); i++) st->iwhich_cgp[CGPATOM(i) - 1] = ig;

    /* any(qconn(:,i,:) <= 3) */
    st->qbonded = xcalloc(S.nat_solute, 1);
    for (i = 1; i <= S.nat_solute; i++) {
        int iq, is, any = 0;
        for (iq = 1; iq <= S.nqat && !any; iq++)
            for (is = 1; is <= S.nstates; is++)
                if (QCONN(is, i, iq) <= 3) { any = 1; break; }
        st->qbonded[i - 1] = (unsigned char)any;
    }

    /* solvent parameters, simprep.f90:3610-3621 */
    st->chg_solv = xcalloc(S.solv_atom, 8);
    st->aLJ_solv = xcalloc(3 * (size_t)S.solv_atom, 8);
    st->bLJ_solv = xcalloc(3 * (size_t)S.solv_atom, 8);
    if (S.nwat > 0)
        for (j = 1; j <= S.solv_atom; j++) {
            CHG_SOLV(j) = CRG(S.nat_solute + j);
            for (c = 1; c <= 3; c++) {
                ALJ_SOLV(j, c) = AVDW(IAC(S.nat_solute + j), c);
                BLJ_SOLV(j, c) = BVDW(IAC(S.nat_solute + j), c);
            }
        }

    /* precompute_interactions, simprep.f90:2860-2870 (pp handled lazily by pp_lookup) */
    if (S.nwat > 0 && S.nat_solute != 0) pw_int_comp(st);
    if (S.nat_solute != 0 && S.nqat != 0) qp_int_comp(st);
    st->nbqq_pair = xcalloc(S.nstates ? S.nstates : 1, sizeof(int));
    st->nbqqp_pair = xcalloc(S.nstates ? S.nstates : 1, sizeof(int));
    if (S.nqat != 0) {
        /* make_nbqqlist, nonbondene.f90:729-745 */
        st->nbqq_max = nbqq_count(st);
        st->nbqq = xcalloc((size_t)st->nbqq_max * S.nstates, sizeof(nbq_t));
        st->nbqqp = xcalloc((size_t)st->nbqq_max * S.nstates, sizeof(nbqp_t));
        if (S.nat_solute == 0) st->qp_precomp = xcalloc(1, sizeof(precomp_t));
        qq_int_comp(st);
        nbqqlist(st);
    }
    if (S.nwat > 0 && S.nqat != 0) qw_int_comp(st);
    if (S.nwat > 0) ww_int_comp(st);

    st->lrf = xcalloc(S.ncgp, sizeof(lrf_t));
    st->boxlength.x = st->boxlength.y = st->boxlength.z = 0;
    st->inv_boxl = st->boxlength;
    return st;
}

void qo_destroy(qo_state *st) {
    if (!st) return;
    free(st->cgp); free(st->cgpatom); free(st->excl); free(st->iqatom); free(st->iqseq); free(st->iac);
    free(st->crg); free(st->iaclib); free(st->ljcod); free(st->listex); free(st->list14);
    free(st->listexlong); free(st->list14long); free(st->qcrg); free(st->qiac); free(st->qavdw);
    free(st->qbvdw); free(st->sc_lookup); free(st->iqexpnb); free(st->jqexpnb); free(st->el_scale_iq);
    free(st->el_scale_jq); free(st->el_scale); free(st->qconn); free(st->iwhich_cgp); free(st->qbonded);
    free(st->chg_solv); free(st->aLJ_solv); free(st->bLJ_solv); free(st->pw_precomp); free(st->ww_precomp);
    free(st->qp_precomp); free(st->qw_precomp); free(st->qq_precomp); free(st->nbpp.p); free(st->nbpw.p);
    free(st->nbww.p); free(st->nbpp_cgp.p); free(st->nbpw_cgp.p); free(st->nbqp_cgp.p); free(st->nbqp);
    free(st->nbqw); free(st->nbqq); free(st->nbqqp); free(st->nbqq_pair); free(st->nbqqp_pair); free(st->lrf);
    free(st);
}

void qo_update_box(qo_state *st, const double boxlength[3], const double inv_boxl[3]) {
    st->boxlength.x = boxlength[0]; st->boxlength.y = boxlength[1]; st->boxlength.z = boxlength[2];
    st->inv_boxl.x = inv_boxl[0]; st->inv_boxl.y = inv_boxl[1]; st->inv_boxl.z = inv_boxl[2];
}

/* list growth: stands in for reallocate_nonbondlist_* (qalloc.f90:370-605) */
static nb_t *nb_push(nb_list *l) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap + l->cap / 2 : 4096;
        l->p = realloc(l->p, (size_t)l->cap * sizeof(nb_t));
        if (!l->p) abort();
    }
    return &l->p[l->n++];
}
static cgp_pair_t *cgp_push(cgp_list *l) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap + l->cap / 2 : 1024;
        l->p = realloc(l->p, (size_t)l->cap * sizeof(cgp_pair_t));
        if (!l->p) abort();
    }
    return &l->p[l->n++];
}
static void nbq_reserve(nbqp_t **p, int64_t *cap, int64_t need, int nstates) {
    if (need > *cap) {
        *cap = need + need / 2 + 1024;
        *p = realloc(*p, (size_t)(*cap) * nstates * sizeof(nbqp_t));
        if (!*p) abort();
    }
}

/* periodic shift as the builders write it: boxlength*q_nint(shift*inv_boxl) */
static inline vec3 box_shift(const qo_state *st, vec3 shift) {
    return v_mul(st->boxlength, q_nint(v_mul(shift, st->inv_boxl)));
}

/* --------------------------------------------------------------------- LRF */

/* cgp_centers, nonbondene.f90:43-78 */
static void cgp_centers(qo_state *st, const double *x) {
    int ig, i;
    for (ig = 1; ig <= S.ncgp; ig++) {
        lrf_t *l = &st->lrf[ig - 1];
        memset(l, 0, sizeof *l);
        double n = (double)(CGP_LAST(ig) - CGP_FIRST(ig) + 1);
        for (i = CGP_FIRST(ig); i <= CGP_LAST(ig); i++) l->cgp_cent = v_add(l->cgp_cent, X(x, CGPATOM(i)));
        /* q_realdiv (math.f90:321): a true division per component, not a multiply by 1/n */
        l->cgp_cent.x = l->cgp_cent.x / n;
        l->cgp_cent.y = l->cgp_cent.y / n;
        l->cgp_cent.z = l->cgp_cent.z / n;
    }
}

/* lrf_update, nonbondene.f90:628-725 */
static void lrf_update(qo_state *st, const double *x, int group1, int group2) {
    int ia, i, n, j;
    vec3 shift = {0, 0, 0};
    lrf_t *l = &st->lrf[group2 - 1];
    if (S.use_PBC) {
        int i_sw = CGP_ISWITCH(group1);
        shift = v_sub(X(x, i_sw), l->cgp_cent);
        shift = box_shift(st, shift);
    }
    for (ia = CGP_FIRST(group1); ia <= CGP_LAST(group1); ia++) {
        double r2, field0, field1, field2, dr[3];
        vec3 drv, t1;
        i = CGPATOM(ia);
        if (IQATOM(i) != 0) continue;
        drv = v_sub(v_sub(X(x, i), l->cgp_cent), shift);
        r2 = qvec_square(drv);
        field0 = CRG(i) / (r2 * sqrt(r2));
        field1 = 3.0 * field0 / r2;
        field2 = -field1 / r2;
        l->phi0 = l->phi0 + field0 * r2;
        l->phi1 = v_sub(l->phi1, v_scale(drv, field0));
        t1 = v_scale(drv, field1);
        l->phi2[0] = v_add(l->phi2[0], v_scale(drv, t1.x));
        l->phi2[1] = v_add(l->phi2[1], v_scale(drv, t1.y));
        l->phi2[2] = v_add(l->phi2[2], v_scale(drv, t1.z));
        l->phi2[0].x -= field0;
        l->phi2[1].y -= field0;
        l->phi2[2].z -= field0;
        dr[0] = drv.x; dr[1] = drv.y; dr[2] = drv.z;
        for (n = 0; n < 3; n++)
            for (j = 0; j < 3; j++) {
                /* tmp(k) = tmp(n)*tmp(j) (= dr_n*dr_j on every component), tmp(l) = ((tmp(k)*dr)*5)*field2 */
                double t = dr[n] * dr[j];
                vec3 tl = v_scale(v_scale(v_scale(drv, t), 5.0), field2);
                l->phi3[j + n * 3] = v_add(l->phi3[j + n * 3], tl);
            }
        t1 = v_scale(drv, r2);
        l->phi3[0].x -= field2 * (3.0 * t1.x);
        l->phi3[0].y -= field2 * (t1.y);
        l->phi3[0].z -= field2 * (t1.z);
        l->phi3[1].x -= field2 * (t1.y);
        l->phi3[1].y -= field2 * (t1.x);
        l->phi3[2].x -= field2 * (t1.z);
        l->phi3[2].z -= field2 * (t1.x);
        l->phi3[3].x -= field2 * (t1.y);
        l->phi3[3].y -= field2 * (t1.x);
        l->phi3[4].x -= field2 * (t1.x);
        l->phi3[4].y -= field2 * (3.0 * t1.y);
        l->phi3[4].z -= field2 * (t1.z);
        l->phi3[5].y -= field2 * (t1.z);
        l->phi3[5].z -= field2 * (t1.y);
        l->phi3[6].x -= field2 * (t1.z);
        l->phi3[6].z -= field2 * (t1.x);
        l->phi3[7].y -= field2 * (t1.z);
        l->phi3[7].z -= field2 * (t1.y);
        l->phi3[8].x -= field2 * (t1.x);
        l->phi3[8].y -= field2 * (t1.y);
        l->phi3[8].z -= field2 * (3.0 * t1.z);
    }
}

/* lrf_taylor, nonbondene.f90:507-571 (natom range = all atoms; sharded callers mask by range) */
static void lrf_taylor(qo_state *st, const double *x, double *d, double *LRF_loc, int a_start, int a_end) {
    int i, ic;
    for (i = a_start; i <= a_end; i++) {
        if ((S.use_PBC && IQATOM(i) == 0) || (!EXCL(i) && IQATOM(i) == 0)) {
            const lrf_t *l;
            vec3 dr, df, tmp;
            double Vij;
            ic = st->iwhich_cgp[i - 1];
            l = &st->lrf[ic - 1];
            dr = v_sub(l->cgp_cent, X(x, i));
            tmp.x = v_dot(dr, l->phi2[0]);
            tmp.y = v_dot(dr, l->phi2[1]);
            tmp.z = v_dot(dr, l->phi2[2]);
            Vij = l->phi0 + v_dot(dr, l->phi1) + 0.5 * v_dot(dr, tmp);
            *LRF_loc = *LRF_loc + 0.5 * CRG(i) * Vij;
            tmp.x = v_dot(dr, l->phi3[0]);
            tmp.y = v_dot(dr, l->phi3[1]);
            tmp.z = v_dot(dr, l->phi3[2]);
            df.x = l->phi1.x + v_dot(dr, l->phi2[0]) + 0.5 * v_dot(dr, tmp);
            tmp.x = v_dot(dr, l->phi3[3]);
            tmp.y = v_dot(dr, l->phi3[4]);
            tmp.z = v_dot(dr, l->phi3[5]);
            df.y = l->phi1.y + v_dot(dr, l->phi2[1]) + 0.5 * v_dot(dr, tmp);
            tmp.x = v_dot(dr, l->phi3[6]);
            tmp.y = v_dot(dr, l->phi3[7]);
            tmp.z = v_dot(dr, l->phi3[8]);
            df.z = l->phi1.z + v_dot(dr, l->phi2[2]) + 0.5 * v_dot(dr, tmp);
            d_sub(d, i, v_scale(df, CRG(i)));
        }
    }
}

/* ------------------------------------------------------------ list builders */

static inline int checkerboard_skip(int ig, int jg) {
    /* e.g. nonbondene.f90:1855-1857 */
    return ((ig > jg) && ((ig + jg) % 2 == 0)) || ((ig < jg) && ((ig + jg) % 2 == 1));
}

/*
 * nbpplist (L1520), nbpplist_lrf (L1777), nbpplist_box (L1643), nbpplist_box_lrf (L1910):
 * switching-atom solute-solute lists; lrf!=0 adds the LRF branch.
 */
static void nbpplist_any(qo_state *st, const double *x, double Rcut, double RLRF, int lrf) {
    int ig, jgr, ia, ja, i, j, is;
    double rcut2 = Rcut, RcLRF2 = RLRF, r2;
    st->nbpp.n = 0;
    st->nbpp_cgp.n = 0;
    for (ig = S.pp_start; ig <= S.pp_end; ig++) {
        is = CGP_ISWITCH(ig);
        if (!S.use_PBC && EXCL(is)) continue;
        for (jgr = 1; jgr <= S.ncgp_solute; jgr++) {
            int jsw;
            if (checkerboard_skip(ig, jgr)) continue;
            jsw = CGP_ISWITCH(jgr);
            if (!S.use_PBC) {
                if (EXCL(jsw)) continue;
                r2 = q_dist4(X(x, is), X(x, jsw));
            } else {
                vec3 shift = v_sub(X(x, is), X(x, jsw));
                r2 = q_dist4(shift, box_shift(st, shift));
            }
            if (r2 <= rcut2) {
                if (S.use_PBC) {
                    cgp_pair_t *c = cgp_push(&st->nbpp_cgp);
                    c->i = is; c->j = jsw;
                }
                for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
                    i = CGPATOM(ia);
                    if (IQATOM(i) != 0) continue;
                    for (ja = CGP_FIRST(jgr); ja <= CGP_LAST(jgr); ja++) {
                        precomp_t p;
                        nb_t *e;
                        j = CGPATOM(ja);
                        if (IQATOM(j) != 0) continue;
                        if (ig == jgr && i >= j) continue;
                        if (!pp_lookup(st, i, j, &p)) continue; /* pp_map == 0 */
                        if (!p.set) continue;
                        e = nb_push(&st->nbpp);
                        e->i = i; e->j = j; e->vdWA = p.vdWA; e->vdWB = p.vdWB; e->elec = p.elec;
                        e->cgp_pair = (int)st->nbpp_cgp.n;
                    }
                }
            } else if (lrf) {
                /* sphere L1891: r2<=RcLRF2; box L2035: (r2<=RcLRF2).or.(RLRF.eq.-one) with RLRF the squared argument */
                if ((r2 <= RcLRF2) || (S.use_PBC && RLRF == -1.0)) {
                    lrf_update(st, x, ig, jgr);
                    lrf_update(st, x, jgr, ig);
                }
            }
        }
    }
}

/* nbpwlist (L2622), nbpwlist_lrf (L2842), nbpwlist_box (L2730), nbpwlist_box_lrf (L2957) */
static void nbpwlist_any(qo_state *st, const double *x, double Rcut, double RLRF, int lrf) {
    int ig, jgr, ia, i, j, is, ja, jg_cgp;
    double rcut2 = Rcut, RcLRF2 = RLRF, r2;
    st->nbpw.n = 0;
    st->nbpw_cgp.n = 0;
    for (ig = S.pw_start; ig <= S.pw_end; ig++) {
        is = CGP_ISWITCH(ig);
        if (!S.use_PBC && EXCL(is)) continue;
        for (jgr = 1; jgr <= S.nwat; jgr++) {
            ja = S.nat_solute + S.solv_atom * jgr - (S.solv_atom - 1);
            jg_cgp = st->iwhich_cgp[ja - 1];
            if (!S.use_PBC) {
                if (EXCL(ja)) continue;
                r2 = q_dist4(X(x, is), X(x, ja));
            } else {
                vec3 shift = v_sub(X(x, is), X(x, ja));
                r2 = q_dist4(shift, box_shift(st, shift));
            }
            if (r2 <= rcut2) {
                if (S.use_PBC) {
                    cgp_pair_t *c = cgp_push(&st->nbpw_cgp);
                    c->i = is; c->j = ja;
                }
                for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
                    i = CGPATOM(ia);
                    if (IQATOM(i) != 0) continue;
                    for (j = 1; j <= S.solv_atom; j++) {
                        nb_t *e = nb_push(&st->nbpw);
                        precomp_t *p = &PW_PRECOMP(i, j);
                        e->i = i;
                        e->j = S.nat_solute + (S.solv_atom * jgr) - S.solv_atom + j;
                        e->vdWA = p->vdWA; e->vdWB = p->vdWB; e->elec = p->elec;
                        e->cgp_pair = (int)st->nbpw_cgp.n;
                    }
                }
            } else if (lrf) {
                /* sphere L2938; box L3066 compares the GLOBAL RcLRF with -one */
                if ((r2 <= RcLRF2) || (S.use_PBC && st->RcLRF_global == -1.0)) {
                    lrf_update(st, x, ig, jg_cgp);
                    lrf_update(st, x, jg_cgp, ig);
                }
            }
        }
    }
}

/* nbwwlist (L4079), nbwwlist_lrf (L4274), nbwwlist_box (L4174), nbwwlist_box_lrf (L4397) */
static void nbwwlist_any(qo_state *st, const double *x, double Rcut, double RLRF, int lrf) {
    int iw, jwr, is, ja, ig, jg, la, ka;
    double rcut2 = Rcut, RcLRF2 = RLRF, r2;
    st->nbww.n = 0;
    for (iw = S.ww_start; iw <= S.ww_end; iw++) {
        is = S.nat_solute + S.solv_atom * iw - (S.solv_atom - 1);
        if (!S.use_PBC && EXCL(is)) continue;
        ig = st->iwhich_cgp[is - 1];
        for (jwr = 1; jwr <= S.nwat; jwr++) {
            ja = S.nat_solute + S.solv_atom * jwr - (S.solv_atom - 1);
            if (!S.use_PBC && EXCL(ja)) continue;
            jg = st->iwhich_cgp[ja - 1];
            if (checkerboard_skip(iw, jwr) || iw == jwr) continue;
            if (!S.use_PBC) r2 = q_dist4(X(x, is), X(x, ja));
            else {
                vec3 shift = v_sub(X(x, is), X(x, ja));
                r2 = q_dist4(shift, box_shift(st, shift));
            }
            if (r2 <= rcut2) {
                for (la = 1; la <= S.solv_atom; la++)
                    for (ka = 1; ka <= S.solv_atom; ka++) {
                        nb_t *e = nb_push(&st->nbww);
                        precomp_t *p = &WW_PRECOMP(la, ka);
                        e->i = S.nat_solute + S.solv_atom * iw - (S.solv_atom - la);
                        e->j = S.nat_solute + S.solv_atom * jwr - (S.solv_atom - ka);
                        e->elec = p->elec; e->vdWA = p->vdWA; e->vdWB = p->vdWB; e->cgp_pair = 0;
                    }
            } else if (lrf) {
                if ((r2 <= RcLRF2) || (S.use_PBC && st->RcLRF_global == -1.0)) {
                    lrf_update(st, x, ig, jg);
                    lrf_update(st, x, jg, ig);
                }
            }
        }
    }
}

/*
 * Any-atom charge-group cut-offs (iuse_switch_atom == 0).
 * nbpplis2 (L950), nbpplis2_lrf (L1369), nbpplis2_box (L1091), nbpplis2_box_lrf (L1220):
 * a group pair is inside when ANY atom pair (Q-atoms included) is within the cut-off.  Sphere+LRF: the LRF
 * test uses r2 of the LAST atom pair examined (L1453-1468, L1499).  Box: the first pair found inside becomes
 * the pair's shift reference (L1165-1173); LRF when any examined pair was within RcLRF2 or the squared
 * argument equals -1 (L1283).
 */
static void nbpplis2_any(qo_state *st, const double *x, double Rcut, double RLRF, int lrf) {
    int ig, jgr, ia, ja, i, j;
    double rcut2 = Rcut, RcLRF2 = RLRF, r2;
    st->nbpp.n = 0;
    st->nbpp_cgp.n = 0;
    for (ig = S.pp_start; ig <= S.pp_end; ig++) {
        if (!S.use_PBC && EXCL(CGP_ISWITCH(ig))) continue;
        for (jgr = 1; jgr <= S.ncgp_solute; jgr++) {
            int inside = 0, inside_LRF = 0;
            if (checkerboard_skip(ig, jgr)) continue;
            if (!S.use_PBC && EXCL(CGP_ISWITCH(jgr))) continue;
            r2 = 1.0;
            ia = CGP_FIRST(ig);
            while (ia <= CGP_LAST(ig) && !inside) {
                i = CGPATOM(ia);
                ja = CGP_FIRST(jgr);
                while (ja <= CGP_LAST(jgr) && !inside) {
                    j = CGPATOM(ja);
                    if (!S.use_PBC) r2 = q_dist4(X(x, i), X(x, j));
                    else {
                        vec3 shift = v_sub(X(x, i), X(x, j));
                        r2 = q_dist4(shift, box_shift(st, shift));
                    }
                    if (r2 <= rcut2) {
                        inside = 1;
                        if (S.use_PBC) {
                            cgp_pair_t *c = cgp_push(&st->nbpp_cgp);
                            c->i = i; c->j = j;
                        }
                    } else if (S.use_PBC && ((r2 <= RcLRF2) || (RLRF == -1.0))) inside_LRF = 1;
                    ja++;
                }
                ia++;
            }
            if (inside) {
                for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
                    i = CGPATOM(ia);
                    if (IQATOM(i) != 0) continue;
                    for (ja = CGP_FIRST(jgr); ja <= CGP_LAST(jgr); ja++) {
                        precomp_t p;
                        nb_t *e;
                        j = CGPATOM(ja);
                        if (IQATOM(j) != 0) continue;
                        if (ig == jgr && i >= j) continue;
                        if (!pp_lookup(st, i, j, &p)) continue;
                        if (!p.set) continue;
                        e = nb_push(&st->nbpp);
                        e->i = i; e->j = j; e->vdWA = p.vdWA; e->vdWB = p.vdWB; e->elec = p.elec;
                        e->cgp_pair = (int)st->nbpp_cgp.n;
                    }
                }
            } else if (lrf) {
                if (S.use_PBC ? inside_LRF : (r2 <= RcLRF2)) {
                    lrf_update(st, x, ig, jgr);
                    lrf_update(st, x, jgr, ig);
                }
            }
        }
    }
}

/* nbpwlis2 (L2137), nbpwlis2_lrf (L2493), nbpwlis2_box (L2247), nbpwlis2_box_lrf (L2359): any solute atom of
 * the group against the water's first atom */
static void nbpwlis2_any(qo_state *st, const double *x, double Rcut, double RLRF, int lrf) {
    int ig, jgr, ia, i, j, ja, jg_cgp;
    double rcut2 = Rcut, RcLRF2 = RLRF, r2;
    st->nbpw.n = 0;
    st->nbpw_cgp.n = 0;
    for (ig = S.pw_start; ig <= S.pw_end; ig++) {
        if (!S.use_PBC && EXCL(CGP_ISWITCH(ig))) continue;
        for (jgr = 1; jgr <= S.nwat; jgr++) {
            int inside = 0, inside_LRF = 0;
            ja = S.nat_solute + S.solv_atom * jgr - (S.solv_atom - 1);
            if (!S.use_PBC && EXCL(ja)) continue;
            jg_cgp = st->iwhich_cgp[ja - 1];
            r2 = 1.0;
            ia = CGP_FIRST(ig);
            while (ia <= CGP_LAST(ig) && !inside) {
                i = CGPATOM(ia);
                if (!S.use_PBC) r2 = q_dist4(X(x, i), X(x, ja));
                else {
                    vec3 shift = v_sub(X(x, i), X(x, ja));
                    r2 = q_dist4(shift, box_shift(st, shift));
                }
                if (r2 <= rcut2) {
                    inside = 1;
                    if (S.use_PBC) {
                        cgp_pair_t *c = cgp_push(&st->nbpw_cgp);
                        c->i = i; c->j = ja;
                    }
                } else if (S.use_PBC && ((r2 <= RcLRF2) || (RLRF == -1.0))) inside_LRF = 1;
                ia++;
            }
            if (inside) {
                for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
                    i = CGPATOM(ia);
                    if (IQATOM(i) != 0) continue;
                    for (j = 1; j <= S.solv_atom; j++) {
                        nb_t *e = nb_push(&st->nbpw);
                        precomp_t *p = &PW_PRECOMP(i, j);
                        e->i = i;
                        e->j = S.nat_solute + (S.solv_atom * jgr) - S.solv_atom + j;
                        e->vdWA = p->vdWA; e->vdWB = p->vdWB; e->elec = p->elec;
                        e->cgp_pair = (int)st->nbpw_cgp.n;
                    }
                }
            } else if (lrf) {
                if (S.use_PBC ? inside_LRF : (r2 <= RcLRF2)) {
                    lrf_update(st, x, ig, jg_cgp);
                    lrf_update(st, x, jg_cgp, ig);
                }
            }
        }
    }
}

static void nbqp_append(qo_state *st, int i) {
    int iq, is;
    nbq_reserve(&st->nbqp, &st->nbqp_cap, st->nbqp_pair + S.nqat, S.nstates);
    for (iq = 1; iq <= S.nqat; iq++) {
        st->nbqp_pair++;
        for (is = 1; is <= S.nstates; is++) {
            nbqp_t *e = &NBQP(st->nbqp_pair, is);
            precomp_t *p = &QP_PRECOMP(i, iq, is);
            e->i = iq; e->j = i; e->vdWA = p->vdWA; e->vdWB = p->vdWB; e->elec = p->elec; e->score = p->score;
            e->cgp_pair = (int)st->nbqp_cgp.n;
        }
    }
}

/* nbqplist, nonbondene.f90:3647-3742 */
static void nbqplist(qo_state *st, const double *x, double Rcut) {
    int ig, ia, i;
    double rcut2, r2;
    vec3 xpcent = {S.xpcent[0], S.xpcent[1], S.xpcent[2]};
    if (st->qp_list_done && (Rcut > S.rexcl_o * S.rexcl_o)) return;
    if (S.nqat == 0) return;
    st->nbqp_pair = 0;
    rcut2 = Rcut;
    for (ig = S.qp_start; ig <= S.qp_end; ig++) {
        ia = CGP_ISWITCH(ig);
        if (EXCL(ia)) continue;
        r2 = q_dist4(X(x, ia), xpcent);
        if (r2 > rcut2) continue;
        for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
            i = CGPATOM(ia);
            if (st->qbonded[i - 1]) continue;
            nbqp_append(st, i);
        }
    }
    st->qp_list_done = 1;
}

/* nbqplis2, nonbondene.f90:3408-3523: any atom of the group within Rcq of xpcent */
static void nbqplis2(qo_state *st, const double *x, double Rcut) {
    int ig, ia, i;
    double rcut2, r2;
    vec3 xpcent = {S.xpcent[0], S.xpcent[1], S.xpcent[2]};
    if (st->qp_list_done && (Rcut > S.rexcl_o * S.rexcl_o)) return;
    st->nbqp_pair = 0;
    rcut2 = Rcut;
    if (S.nqat == 0) return;
    for (ig = S.qp_start; ig <= S.qp_end; ig++) {
        int inside = 0;
        if (EXCL(CGP_ISWITCH(ig))) continue;
        ia = CGP_FIRST(ig);
        while (ia <= CGP_LAST(ig) && !inside) {
            i = CGPATOM(ia);
            r2 = q_dist4(X(x, i), xpcent);
            if (r2 <= rcut2) inside = 1;
            ia++;
        }
        if (!inside) continue;
        for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
            i = CGPATOM(ia);
            if (st->qbonded[i - 1]) continue;
            nbqp_append(st, i);
        }
    }
    st->qp_list_done = 1;
}

/*
 * nbqplist_box, nonbondene.f90:3747-3859.  The reference neither resets nbqp_cgp_pair nor
 * stores %cgp_pair here (SURVEY quirk iii) so nonbond_qp_box reads an undefined shift;
 * the oracle follows the sibling nbqplis2_box (L3562-3565, L3631) which does both.  For
 * Rq<0 no group pair is ever registered and the reference indexes nbqp_cgp(0): the oracle
 * defines that shift as zero (cgp_pair == 0 -> zero shift in nonbond_qp_box below).
 */
static void nbqplist_box(qo_state *st, const double *x, double Rcut, double Rq) {
    int ig, ia, i, inside, ig_atom;
    double rcut2, r2;
    if (st->qp_list_done && (S.use_PBC && Rq < 0.0)) return;
    if (S.nqat == 0) return;
    st->nbqp_pair = 0;
    st->nbqp_cgp.n = 0;
    rcut2 = Rcut;
    for (ig = S.qp_start; ig <= S.qp_end; ig++) {
        inside = (Rq < 0.0) ? 1 : 0;
        ig_atom = CGP_FIRST(ig);
        while (ig_atom <= CGP_LAST(ig) && inside == 0) {
            vec3 shift;
            i = CGPATOM(ig_atom);
            shift = v_sub(X(x, i), X(x, S.qswitch));
            r2 = q_dist4(shift, box_shift(st, shift));
            if (r2 <= rcut2) {
                cgp_pair_t *c;
                inside = 1;
                c = cgp_push(&st->nbqp_cgp);
                c->i = i; c->j = S.qswitch;
            }
            ig_atom++;
        }
        if (inside == 0) continue;
        for (ia = CGP_FIRST(ig); ia <= CGP_LAST(ig); ia++) {
            i = CGPATOM(ia);
            if (st->qbonded[i - 1]) continue;
            nbqp_append(st, i);
        }
    }
    st->qp_list_done = 1;
}

static void nbqw_append(qo_state *st, int ia) {
    int iq, is, k;
    nbq_reserve(&st->nbqw, &st->nbqw_cap, st->nbqw_pair + (int64_t)S.nqat * S.solv_atom, S.nstates);
    for (iq = 1; iq <= S.nqat; iq++)
        for (is = 1; is <= S.solv_atom; is++) {
            st->nbqw_pair++;
            for (k = 1; k <= S.nstates; k++) {
                nbqp_t *e = &NBQW(st->nbqw_pair, k);
                precomp_t *p = &QW_PRECOMP(iq, is, k);
                e->i = iq; e->j = ia + (is - 1);
                e->elec = p->elec; e->vdWA = p->vdWA; e->vdWB = p->vdWB; e->score = p->score; e->cgp_pair = 0;
            }
        }
}

/* nbqwlist, nonbondene.f90:3862-3944 */
static void nbqwlist(qo_state *st, const double *x, double Rcut) {
    int ig, ia;
    double rcut2, r2;
    vec3 xpcent = {S.xpcent[0], S.xpcent[1], S.xpcent[2]};
    if (st->qw_list_done && (Rcut > S.rexcl_o * S.rexcl_o)) return;
    if (S.nqat == 0) return;
    st->nbqw_pair = 0;
    rcut2 = Rcut;
    for (ig = S.qw_start; ig <= S.qw_end; ig++) {
        ia = S.nat_solute + S.solv_atom * ig - (S.solv_atom - 1);
        if (EXCL(ia)) continue;
        r2 = q_dist4(X(x, ia), xpcent);
        if (r2 <= rcut2) nbqw_append(st, ia);
    }
    st->qw_list_done = 1;
}

/* nbqwlist_box, nonbondene.f90:3949-4025 */
static void nbqwlist_box(qo_state *st, const double *x, double Rcut, double Rq) {
    int ig, ia;
    double rcut2, r2;
    if (st->qw_list_done && Rq < 0.0) return;
    if (S.nqat == 0) return;
    st->nbqw_pair = 0;
    rcut2 = Rcut;
    for (ig = S.qw_start; ig <= S.qw_end; ig++) {
        ia = S.nat_solute + S.solv_atom * ig - (S.solv_atom - 1);
        if (Rq > 0.0) {
            vec3 shift = v_sub(X(x, ia), X(x, S.qswitch));
            r2 = q_dist4(shift, box_shift(st, shift));
        } else r2 = 1.0;
        if ((r2 <= rcut2) || (Rq < 0.0)) nbqw_append(st, ia);
    }
    st->qw_list_done = 1;
}

/* make_pair_lists, nonbondene.f90:749-837 */
int qo_make_pair_lists(qo_state *st, const double *x, double Rq, double Rcq2, double RcLRF2, double Rcpp2,
                       double Rcpw2, double Rcww2, double RcLRF, int64_t counts_out[8]) {
    const int sw = S.iuse_switch_atom == 1, lrf = S.use_LRF != 0;
    st->RcLRF_global = RcLRF;
    if (lrf) cgp_centers(st, x);
    if (sw) {
        nbpplist_any(st, x, Rcpp2, RcLRF2, lrf);
        nbpwlist_any(st, x, Rcpw2, RcLRF2, lrf);
    } else {
        nbpplis2_any(st, x, Rcpp2, RcLRF2, lrf);
        nbpwlis2_any(st, x, Rcpw2, RcLRF2, lrf);
    }
    if (S.use_PBC) {
        /* nbqplist_box and nbqplis2_box are the same any-atom scan; they differ only in the %cgp_pair
           bookkeeping that the oracle defines as in nbqplis2_box (see nbqplist_box above) */
        nbqplist_box(st, x, Rcq2, Rq);
        nbwwlist_any(st, x, Rcww2, RcLRF2, lrf);
        nbqwlist_box(st, x, Rcq2, Rq);
    } else {
        if (sw) nbqplist(st, x, Rcq2); else nbqplis2(st, x, Rcq2);
        nbwwlist_any(st, x, Rcww2, RcLRF2, lrf);
        nbqwlist(st, x, Rcq2);
    }
    if (counts_out) {
        counts_out[0] = st->nbpp.n; counts_out[1] = st->nbpw.n; counts_out[2] = st->nbww.n;
        counts_out[3] = st->nbqp_pair; counts_out[4] = st->nbqw_pair;
        counts_out[5] = st->nbpp_cgp.n; counts_out[6] = st->nbpw_cgp.n; counts_out[7] = st->nbqp_cgp.n;
    }
    return 0;
}

/* -------------------------------------------------------- pair energy functions */

/* q_dist2 (math.f90:226-238) + nbe (nonbonded.f90:45-75); vec = x(j)-x(i) [+shift] */
static eneret_t nbe_vec(const nb_t *nb, vec3 vec) {
    eneret_t e;
    double r2 = 1.0 / qvec_square(vec);
    double r = sqrt(r2);
    double r6 = r2 * r2 * r2;
    double r12 = r6 * r6;
    e.Vel = nb->elec * r;
    e.V_a = nb->vdWA * r12;
    e.V_b = nb->vdWB * r6;
    e.dv = r2 * (-e.Vel - 12.0 * e.V_a + 6.0 * e.V_b);
    e.vec = vec;
    return e;
}
/* q_dist (math.f90:214-224) + nbe_spc/nbe_spcb (nonbonded.f90:247-291) */
static eneret_t nbe_spc_vec(const nb_t *nb, vec3 vec) {
    eneret_t e;
    double r2 = 1.0 / qvec_square(vec);
    double r = sqrt(r2);
    e.vec = vec;
    e.Vel = nb->elec * r;
    e.V_a = e.V_b = 0;
    e.dv = r2 * (-e.Vel);
    return e;
}
/* q_dist3 (math.f90:240-252) + nbe_qx (nonbonded.f90:200-222) */
typedef struct { double r2, r, r6; vec3 vec; } dist3_t;
static dist3_t q_dist3_vec(vec3 vec) {
    dist3_t d;
    d.r2 = 1.0 / qvec_square(vec);
    d.r = sqrt(d.r2);
    d.r6 = d.r2 * d.r2 * d.r2;
    d.vec = vec;
    return d;
}
static eneret_t nbe_qx(const nbqp_t *nb, double lambda, const dist3_t *dist) {
    eneret_t e;
    double r2 = dist->r2, r6_hc = 1.0 / dist->r6, r = dist->r, r6, r12;
    r6 = r6_hc + nb->score;
    r6 = 1.0 / r6;
    r12 = r6 * r6;
    e.Vel = nb->elec * r;
    e.V_a = nb->vdWA * r12;
    e.V_b = nb->vdWB * r6;
    e.dv = r2 * (-e.Vel - (12.0 * e.V_a - 6.0 * e.V_b) * r6 * r6_hc) * lambda;
    e.vec = dist->vec;
    return e;
}
/* nbe_qspc, nonbonded.f90:226-243 */
static eneret_t nbe_qspc(const nbqp_t *nb, double lambda, double r2, double r) {
    eneret_t e;
    e.Vel = nb->elec * r;
    e.V_a = e.V_b = 0;
    e.dv = -r2 * e.Vel * lambda;
    return e;
}
/* nbe_qq, nonbonded.f90:111-150 */
static eneret_t nbe_qq(const nbq_t *nb, double lambda, vec3 vec) {
    eneret_t e;
    dist3_t d = q_dist3_vec(vec);
    double r2 = d.r2, r6_hc = 1.0 / d.r6, r = d.r, r6, r12;
    r6 = r6_hc + nb->score;
    r6 = 1.0 / r6;
    r12 = r6 * r6;
    e.Vel = nb->elec * r;
    e.vec = vec;
    if (nb->soft) {
        e.V_b = 0;
        e.V_a = nb->vdWA * exp(-nb->vdWB / r);
        e.dv = r2 * (-e.Vel - nb->vdWB * e.V_a / r) * lambda;
    } else {
        e.V_a = nb->vdWA * r12;
        e.V_b = nb->vdWB * r6;
        e.dv = r2 * (-e.Vel - (12.0 * e.V_a - 6.0 * e.V_b) * r6 * r6_hc) * lambda;
    }
    return e;
}

/* ------------------------------------------------------------- force loops */

/* nonbond_pp (L4694) / nonbond_pw (L4856): same loop over an NB_TYPE list */
static void nonbond_list(const nb_list *l, const double *x, double *d, double *el, double *vdw) {
    int64_t ip;
    for (ip = 0; ip < l->n; ip++) {
        const nb_t *nb = &l->p[ip];
        eneret_t e = nbe_vec(nb, v_sub(X(x, nb->j), X(x, nb->i)));
        d_sub(d, nb->i, v_scale(e.vec, e.dv));
        d_add(d, nb->j, v_scale(e.vec, e.dv));
        *el = *el + e.Vel;
        *vdw = *vdw + e.V_a - e.V_b;
    }
}
/* nonbond_pp_box (L4763) / nonbond_pw_box (L4920) */
static void nonbond_list_box(qo_state *st, const nb_list *l, cgp_list *cl, const double *x, double *d, double *el,
                             double *vdw) {
    int64_t ip, g;
    for (g = 0; g < cl->n; g++) {
        vec3 shift = v_sub(X(x, cl->p[g].i), X(x, cl->p[g].j));
        cl->p[g].shift = box_shift(st, shift);
    }
    for (ip = 0; ip < l->n; ip++) {
        const nb_t *nb = &l->p[ip];
        /* nbe_b: q_dist2(x(i)-x(j), shift) -> vec = shift - (x(i)-x(j)) */
        vec3 vec = v_sub(cl->p[nb->cgp_pair - 1].shift, v_sub(X(x, nb->i), X(x, nb->j)));
        eneret_t e = nbe_vec(nb, vec);
        d_sub(d, nb->i, v_scale(e.vec, e.dv));
        d_add(d, nb->j, v_scale(e.vec, e.dv));
        *el = *el + e.Vel;
        *vdw = *vdw + e.V_a - e.V_b;
    }
}

/* nonbond_ww (L5509) and nonbond_ww_box (L5576): general solvent */
static void nonbond_ww_general(qo_state *st, const double *x, double *d, double *el, double *vdw) {
    int64_t iw, ia, n2 = (int64_t)S.solv_atom * S.solv_atom;
    for (iw = 0; iw < st->nbww.n; iw += n2) {
        vec3 shift = {0, 0, 0};
        if (S.use_PBC) shift = box_shift(st, v_sub(X(x, st->nbww.p[iw].i), X(x, st->nbww.p[iw].j)));
        for (ia = 0; ia < n2; ia++) {
            const nb_t *nb = &st->nbww.p[iw + ia];
            vec3 vec = S.use_PBC ? v_sub(shift, v_sub(X(x, nb->i), X(x, nb->j))) : v_sub(X(x, nb->j), X(x, nb->i));
            eneret_t e = nbe_vec(nb, vec);
            d_sub(d, nb->i, v_scale(e.vec, e.dv));
            d_add(d, nb->j, v_scale(e.vec, e.dv));
            *el = *el + e.Vel;
            *vdw = *vdw + e.V_a - e.V_b;
        }
    }
}
/* nonbond_ww_spc (L5886) and nonbond_ww_spc_box (L5979) */
static void nonbond_ww_spc(qo_state *st, const double *x, double *d, double *el, double *vdw) {
    int64_t iw, jw, n2 = (int64_t)S.solv_atom * S.solv_atom;
    for (iw = 0; iw < st->nbww.n; iw += n2) {
        const nb_t *nb = &st->nbww.p[iw];
        vec3 shift = {0, 0, 0}, vec;
        eneret_t e;
        if (S.use_PBC) {
            shift = box_shift(st, v_sub(X(x, nb->i), X(x, nb->j)));
            vec = v_sub(shift, v_sub(X(x, nb->i), X(x, nb->j)));
        } else vec = v_sub(X(x, nb->j), X(x, nb->i));
        e = nbe_vec(nb, vec);
        *vdw = *vdw + e.V_a - e.V_b;
        *el = *el + e.Vel;
        d_sub(d, nb->i, v_scale(e.vec, e.dv));
        d_add(d, nb->j, v_scale(e.vec, e.dv));
        for (jw = 1; jw < n2; jw++) {
            nb = &st->nbww.p[iw + jw];
            vec = S.use_PBC ? v_sub(shift, v_sub(X(x, nb->i), X(x, nb->j))) : v_sub(X(x, nb->j), X(x, nb->i));
            e = nbe_spc_vec(nb, vec);
            *el = *el + e.Vel;
            d_sub(d, nb->i, v_scale(e.vec, e.dv));
            d_add(d, nb->j, v_scale(e.vec, e.dv));
        }
    }
}

/* nonbond_qp (L5157) and nonbond_qp_box (L5234) */
static void nonbond_qp(qo_state *st, const double *x, const double *lambda, double *d, double *EQ /*qp el,vdw per state: stride QNB_EQ_STRIDE*/) {
    int64_t ip, g;
    int istate;
    if (S.use_PBC)
        for (g = 0; g < st->nbqp_cgp.n; g++) {
            vec3 shift = v_sub(X(x, st->nbqp_cgp.p[g].i), X(x, S.qswitch));
            st->nbqp_cgp.p[g].shift = box_shift(st, shift);
        }
    for (ip = 1; ip <= st->nbqp_pair; ip++) {
        int iq = NBQP(ip, 1).i, i = IQSEQ(iq), j = NBQP(ip, 1).j;
        dist3_t dist;
        if (S.use_PBC) {
            int group = NBQP(ip, 1).cgp_pair;
            vec3 sh = {0, 0, 0};
            if (group >= 1) sh = st->nbqp_cgp.p[group - 1].shift; /* group 0: see nbqplist_box note */
            dist = q_dist3_vec(v_sub(sh, v_sub(X(x, i), X(x, j))));
        } else dist = q_dist3_vec(v_sub(X(x, j), X(x, i)));
        for (istate = 1; istate <= S.nstates; istate++) {
            eneret_t e = nbe_qx(&NBQP(ip, istate), lambda[istate - 1], &dist);
            d_sub(d, i, v_scale(dist.vec, e.dv));
            d_add(d, j, v_scale(dist.vec, e.dv));
            EQ[(istate - 1) * QNB_EQ_STRIDE + 2] += e.Vel;
            EQ[(istate - 1) * QNB_EQ_STRIDE + 3] += e.V_a - e.V_b;
        }
    }
}

/* nonbond_qw (L5339) and nonbond_qw_box (L5422): general solvent */
static void nonbond_qw_general(qo_state *st, const double *x, const double *lambda, double *d, double *EQ) {
    int64_t iw;
    int jw, istate;
    for (iw = 1; iw <= st->nbqw_pair; iw += S.solv_atom) {
        int iq = NBQW(iw, 1).i, i = IQSEQ(iq);
        for (jw = 0; jw <= S.solv_atom - 1; jw++) {
            int64_t jp = iw + jw;
            int j = NBQW(jp, 1).j;
            dist3_t dist;
            if (S.use_PBC) {
                /* L5470-5472: the shift is computed per atom j */
                vec3 shift = box_shift(st, v_sub(X(x, S.qswitch), X(x, j)));
                dist = q_dist3_vec(v_sub(shift, v_sub(X(x, i), X(x, j))));
            } else dist = q_dist3_vec(v_sub(X(x, j), X(x, i)));
            for (istate = 1; istate <= S.nstates; istate++) {
                eneret_t e = nbe_qx(&NBQW(jp, istate), lambda[istate - 1], &dist);
                d_sub(d, i, v_scale(dist.vec, e.dv));
                d_add(d, j, v_scale(dist.vec, e.dv));
                EQ[(istate - 1) * QNB_EQ_STRIDE + 4] += e.Vel;
                EQ[(istate - 1) * QNB_EQ_STRIDE + 5] += e.V_a - e.V_b;
            }
        }
    }
}

/* nonbond_qw_spc (L5658) and nonbond_qw_spc_box (L5771) */
static void nonbond_qw_spc(qo_state *st, const double *x, const double *lambda, double *d, double *EQ) {
    int64_t iw;
    int ip, istate;
    for (iw = 1; iw <= st->nbqw_pair; iw += S.solv_atom) {
        int iq = NBQW(iw, 1).i, i = IQSEQ(iq), j = NBQW(iw, 1).j;
        vec3 shift = {0, 0, 0};
        dist3_t dist;
        double dv = 0;
        if (S.use_PBC) {
            shift = box_shift(st, v_sub(X(x, S.qswitch), X(x, j)));
            dist = q_dist3_vec(v_sub(shift, v_sub(X(x, i), X(x, j))));
        } else dist = q_dist3_vec(v_sub(X(x, j), X(x, i)));
        for (istate = 1; istate <= S.nstates; istate++) {
            eneret_t e = nbe_qx(&NBQW(iw, istate), lambda[istate - 1], &dist);
            dv = dv + e.dv;
            EQ[(istate - 1) * QNB_EQ_STRIDE + 4] += e.Vel;
            EQ[(istate - 1) * QNB_EQ_STRIDE + 5] += e.V_a - e.V_b;
        }
        d_sub(d, i, v_scale(dist.vec, dv));
        d_add(d, j, v_scale(dist.vec, dv));
        for (ip = 1; ip <= S.solv_atom - 1; ip++) {
            int64_t jp = iw + ip;
            int ja = j + ip;
            vec3 vec = S.use_PBC ? v_sub(shift, v_sub(X(x, i), X(x, ja))) : v_sub(X(x, ja), X(x, i));
            double r2 = 1.0 / qvec_square(vec), r = sqrt(r2);
            dv = 0;
            for (istate = 1; istate <= S.nstates; istate++) {
                eneret_t e = nbe_qspc(&NBQW(jp, istate), lambda[istate - 1], r2, r);
                EQ[(istate - 1) * QNB_EQ_STRIDE + 4] += e.Vel;
                dv = dv + e.dv;
            }
            d_sub(d, i, v_scale(vec, dv));
            d_add(d, ja, v_scale(vec, dv));
        }
    }
}

/* nonbond_qq, nonbondene.f90:5013-5083 */
static void nonbond_qq(qo_state *st, const double *x, const double *lambda, double *d, double *EQ) {
    int istate, ip;
    for (istate = 1; istate <= S.nstates; istate++)
        for (ip = 1; ip <= st->nbqq_pair[istate - 1]; ip++) {
            const nbq_t *nb = &NBQQ(ip, istate);
            int i = IQSEQ(nb->iq), j = IQSEQ(nb->jq);
            eneret_t e = nbe_qq(nb, lambda[istate - 1], v_sub(X(x, j), X(x, i)));
            d_sub(d, i, v_scale(e.vec, e.dv));
            d_add(d, j, v_scale(e.vec, e.dv));
            EQ[(istate - 1) * QNB_EQ_STRIDE + 0] += e.Vel;
            EQ[(istate - 1) * QNB_EQ_STRIDE + 1] += e.V_a - e.V_b;
        }
}
/* nonbond_qqp, nonbondene.f90:5085-5155: energies go to EQ%qp (potene.f90:177) */
static void nonbond_qqp(qo_state *st, const double *x, const double *lambda, double *d, double *EQ) {
    int istate, ip;
    for (istate = 1; istate <= S.nstates; istate++)
        for (ip = 1; ip <= st->nbqqp_pair[istate - 1]; ip++) {
            const nbqp_t *nb = &NBQQP(ip, istate);
            int i = IQSEQ(nb->i), j = nb->j;
            dist3_t dist = q_dist3_vec(v_sub(X(x, j), X(x, i)));
            eneret_t e = nbe_qx(nb, lambda[istate - 1], &dist);
            d_sub(d, i, v_scale(dist.vec, e.dv));
            d_add(d, j, v_scale(dist.vec, e.dv));
            EQ[(istate - 1) * QNB_EQ_STRIDE + 2] += e.Vel;
            EQ[(istate - 1) * QNB_EQ_STRIDE + 3] += e.V_a - e.V_b;
        }
}

/* pot_energy_nonbonds, potene.f90:320-380, then nonbond_qq/qqp (potene.f90:176-177) */
int qo_nonbond(qo_state *st, const double *x, const double *lambda, int flags, double *d, double *E, double *EQ) {
    int md = (flags & QNB_FLAG_MD) != 0, k;
    int spc = (S.ivdw_rule == QNB_VDW_GEOMETRIC) && (S.solvent_type == QNB_SOLVENT_SPC);
    for (k = 0; k < QNB_E_COUNT; k++) E[k] = 0;
    for (k = 0; k < QNB_EQ_STRIDE * S.nstates; k++) EQ[k] = 0;
    if (S.use_PBC) {
        if (S.natom > S.nat_solute) {
            if (spc) {
                nonbond_qw_spc(st, x, lambda, d, EQ);
                if (md) nonbond_ww_spc(st, x, d, &E[QNB_E_WW_EL], &E[QNB_E_WW_VDW]);
            } else {
                nonbond_qw_general(st, x, lambda, d, EQ);
                if (md) nonbond_ww_general(st, x, d, &E[QNB_E_WW_EL], &E[QNB_E_WW_VDW]);
            }
            if (md) nonbond_list_box(st, &st->nbpw, &st->nbpw_cgp, x, d, &E[QNB_E_PW_EL], &E[QNB_E_PW_VDW]);
        }
        if (md) nonbond_list_box(st, &st->nbpp, &st->nbpp_cgp, x, d, &E[QNB_E_PP_EL], &E[QNB_E_PP_VDW]);
        nonbond_qp(st, x, lambda, d, EQ);
    } else {
        if (S.natom > S.nat_solute) {
            if (spc) {
                if (md) nonbond_ww_spc(st, x, d, &E[QNB_E_WW_EL], &E[QNB_E_WW_VDW]);
                nonbond_qw_spc(st, x, lambda, d, EQ);
            } else {
                nonbond_qw_general(st, x, lambda, d, EQ);
                if (md) nonbond_ww_general(st, x, d, &E[QNB_E_WW_EL], &E[QNB_E_WW_VDW]);
            }
            if (md) nonbond_list(&st->nbpw, x, d, &E[QNB_E_PW_EL], &E[QNB_E_PW_VDW]);
        }
        if (md) nonbond_list(&st->nbpp, x, d, &E[QNB_E_PP_EL], &E[QNB_E_PP_VDW]);
        nonbond_qp(st, x, lambda, d, EQ);
    }
    if (S.use_LRF && md) {
        /* calculation_assignment%natom: the shard's atom range */
        lrf_taylor(st, x, d, &E[QNB_E_LRF], S.natom_start, S.natom_end);
    }
    if (flags & QNB_FLAG_QQ) {
        nonbond_qq(st, x, lambda, d, EQ);
        nonbond_qqp(st, x, lambda, d, EQ);
    }
    return 0;
}

/* ---------------------------------------------------------------- exports */
int qo_list_count(qo_state *st, int which, int state, int64_t *n) {
    switch (which) {
    case QNB_LIST_PP: *n = st->nbpp.n; return 0;
    case QNB_LIST_PW: *n = st->nbpw.n; return 0;
    case QNB_LIST_WW: *n = st->nbww.n; return 0;
    case QNB_LIST_QP: *n = st->nbqp_pair; return 0;
    case QNB_LIST_QW: *n = st->nbqw_pair; return 0;
    case QNB_LIST_QQ: *n = st->nbqq_pair[state - 1]; return 0;
    case QNB_LIST_QQP: *n = st->nbqqp_pair[state - 1]; return 0;
    }
    return 1;
}

int qo_export_list(qo_state *st, int which, int state, int32_t *ij, double *params, int64_t capacity) {
    int64_t n, k;
    if (qo_list_count(st, which, state, &n)) return 1;
    if (n > capacity) return 2;
    for (k = 0; k < n; k++) {
        double A = 0, B = 0, el = 0, sc = 0;
        int i = 0, j = 0;
        if (which <= QNB_LIST_WW) {
            const nb_list *l = which == QNB_LIST_PP ? &st->nbpp : which == QNB_LIST_PW ? &st->nbpw : &st->nbww;
            i = l->p[k].i; j = l->p[k].j; A = l->p[k].vdWA; B = l->p[k].vdWB; el = l->p[k].elec;
        } else if (which == QNB_LIST_QP || which == QNB_LIST_QW) {
            const nbqp_t *e = which == QNB_LIST_QP ? &NBQP(k + 1, state) : &NBQW(k + 1, state);
            i = e->i; j = e->j; A = e->vdWA; B = e->vdWB; el = e->elec; sc = e->score;
        } else if (which == QNB_LIST_QQ) {
            const nbq_t *e = &NBQQ(k + 1, state);
            i = e->iq; j = e->jq; A = e->vdWA; B = e->vdWB; el = e->elec; sc = e->score;
        } else {
            const nbqp_t *e = &NBQQP(k + 1, state);
            i = e->i; j = e->j; A = e->vdWA; B = e->vdWB; el = e->elec; sc = e->score;
        }
        ij[2 * k] = i; ij[2 * k + 1] = j;
        if (params) { params[4 * k] = A; params[4 * k + 1] = B; params[4 * k + 2] = el; params[4 * k + 3] = sc; }
    }
    return 0;
}

int qo_export_lrf(qo_state *st, double *out) {
    int ig, k;
    for (ig = 0; ig < S.ncgp; ig++) {
        const lrf_t *l = &st->lrf[ig];
        double *o = out + (size_t)QNB_LRF_STRIDE * ig;
        o[0] = l->cgp_cent.x; o[1] = l->cgp_cent.y; o[2] = l->cgp_cent.z;
        o[3] = l->phi0;
        o[4] = l->phi1.x; o[5] = l->phi1.y; o[6] = l->phi1.z;
        for (k = 0; k < 3; k++) { o[7 + 3 * k] = l->phi2[k].x; o[8 + 3 * k] = l->phi2[k].y; o[9 + 3 * k] = l->phi2[k].z; }
        for (k = 0; k < 9; k++) { o[16 + 3 * k] = l->phi3[k].x; o[17 + 3 * k] = l->phi3[k].y; o[18 + 3 * k] = l->phi3[k].z; }
    }
    return 0;
}

/* -------------------------------------------------------------- make_qconn */
typedef struct {
    int nstates, nat_solute, nqat, nbonds_solute, nqbond;
    const int32_t *iqatom, *bnd, *qbnd_ij, *qbnd_cod;
    int32_t *qconn;
} qconn_ctx;
#define QC(is, i, iq) (c->qconn[((is)-1) + (size_t)((i)-1) * c->nstates + (size_t)((iq)-1) * c->nstates * c->nat_solute])

/* find_bonded, nonbondene.f90:3131-3187 */
static void find_bonded(qconn_ctx *c, int origin, int current, int level, int state) {
    int b, newlevel, newcurrent, oq = c->iqatom[origin - 1];
    for (b = 0; b < c->nbonds_solute; b++) {
        int bi = c->bnd[3 * b], bj = c->bnd[3 * b + 1], cod = c->bnd[3 * b + 2];
        if (cod == 0) continue;
        if (bi == current) newcurrent = bj;
        else if (bj == current) newcurrent = bi;
        else continue;
        newlevel = level + 1;
        if (QC(state, newcurrent, oq) > newlevel) {
            QC(state, newcurrent, oq) = newlevel;
            if (newlevel < 4) find_bonded(c, origin, newcurrent, newlevel, state);
        }
    }
    for (b = 0; b < c->nqbond; b++) {
        if (c->qbnd_cod[b + (size_t)(state - 1) * c->nqbond] > 0) {
            int bi = c->qbnd_ij[2 * b], bj = c->qbnd_ij[2 * b + 1];
            if (bi == current) newcurrent = bj;
            else if (bj == current) newcurrent = bi;
            else continue;
            newlevel = level + 1;
            if (QC(state, newcurrent, oq) > newlevel) {
                QC(state, newcurrent, oq) = newlevel;
                if (newlevel < 4) find_bonded(c, origin, newcurrent, newlevel, state);
            }
        }
    }
}

/* make_qconn, nonbondene.f90:3087-3125 */
void qo_make_qconn(int nstates, int nat_solute, int nqat, const int32_t *iqseq, const int32_t *iqatom,
                   int nbonds_solute, const int32_t *bnd, int nqbond, const int32_t *qbnd_ij,
                   const int32_t *qbnd_cod, int nexspec, const int32_t *exspec_ij, const int32_t *exspec_flag,
                   int32_t *qconn) {
    qconn_ctx ctx = {nstates, nat_solute, nqat, nbonds_solute, nqbond, iqatom, bnd, qbnd_ij, qbnd_cod, qconn};
    qconn_ctx *c = &ctx;
    size_t n = (size_t)nstates * nat_solute * nqat, k;
    int iq, is, i;
    for (k = 0; k < n; k++) qconn[k] = 9;
    for (iq = 1; iq <= nqat; iq++)
        for (is = 1; is <= nstates; is++) QC(is, iqseq[iq - 1], iq) = 1;
    for (iq = 1; iq <= nqat; iq++)
        for (is = 1; is <= nstates; is++) {
            i = iqseq[iq - 1];
            find_bonded(c, i, i, 1, is);
        }
    for (i = 0; i < nexspec; i++) {
        int a = exspec_ij[2 * i], b = exspec_ij[2 * i + 1];
        iq = iqatom[a - 1];
        if (iq > 0)
            for (is = 1; is <= nstates; is++)
                if (exspec_flag[i + (size_t)(is - 1) * nexspec]) QC(is, b, iq) = 0;
        iq = iqatom[b - 1];
        if (iq > 0)
            for (is = 1; is <= nstates; is++)
                if (exspec_flag[i + (size_t)(is - 1) * nexspec]) QC(is, a, iq) = 0;
    }
}

/* ------------------------------------------------------- decomposed timing */
typedef struct {
    qo_state *st;
    const double *x, *lambda, *cut;
    int flags, steps, build_each;
    double *d, E[QNB_E_COUNT], *EQ;
    double list_s, force_s;
} worker_t;

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *worker_lists(void *arg) {
    worker_t *w = arg;
    double t0 = now_s();
    qo_make_pair_lists(w->st, w->x, w->cut[0], w->cut[1], w->cut[2], w->cut[3], w->cut[4], w->cut[5], w->cut[6], NULL);
    w->list_s += now_s() - t0;
    return NULL;
}
static void *worker_force(void *arg) {
    worker_t *w = arg;
    int k;
    double t0 = now_s();
    for (k = 0; k < w->steps; k++) {
        if (w->build_each) {
            w->st->qp_list_done = w->st->qw_list_done = 0;
            qo_make_pair_lists(w->st, w->x, w->cut[0], w->cut[1], w->cut[2], w->cut[3], w->cut[4], w->cut[5], w->cut[6], NULL);
        }
        memset(w->d, 0, sizeof(double) * 3 * (size_t)w->st->s.natom);
        qo_nonbond(w->st, w->x, w->lambda, w->flags, w->d, w->E, w->EQ);
    }
    w->force_s += now_s() - t0;
    return NULL;
}

/* contiguous ranges balanced by count (equal shares; distribute_nonbonds without the master discount) */
static void split_range(int n, int nparts, int part, int *start, int *end) {
    int q = n / nparts, r = n % nparts;
    *start = part * q + (part < r ? part : r) + 1;
    *end = *start + q - 1 + (part < r ? 1 : 0);
}

double qo_time_decomposed(const qnb_system *sys, const double *x, const double *lambda, int flags,
                          const double cut[7], const double *box6, int nthreads, int steps, int build_each,
                          double *d_out, double *E_out, double *EQ_out, double *list_seconds) {
    worker_t *w = xcalloc(nthreads, sizeof *w);
    pthread_t *th = xcalloc(nthreads, sizeof *th);
    int t, k;
    double t0, wall, lmax = 0;
    size_t n3 = 3 * (size_t)sys->natom;
    for (t = 0; t < nthreads; t++) {
        qnb_system s = *sys;
        split_range(sys->ncgp_solute, nthreads, t, &s.pp_start, &s.pp_end);
        s.pw_start = s.qp_start = s.pp_start;
        s.pw_end = s.qp_end = s.pp_end;
        split_range(sys->nwat, nthreads, t, &s.ww_start, &s.ww_end);
        s.qw_start = s.ww_start;
        s.qw_end = s.ww_end;
        split_range(sys->natom, nthreads, t, &s.natom_start, &s.natom_end);
        s.is_master = (t == 0);
        w[t].st = qo_create(&s);
        if (!w[t].st) return -1;
        if (box6) qo_update_box(w[t].st, box6, box6 + 3);
        w[t].x = x; w[t].lambda = lambda; w[t].cut = cut; w[t].steps = steps; w[t].build_each = build_each;
        w[t].flags = (t == 0) ? flags : (flags & ~QNB_FLAG_QQ);
        w[t].d = xcalloc(n3, 8);
        w[t].EQ = xcalloc((size_t)QNB_EQ_STRIDE * sys->nstates + 1, 8);
    }
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker_lists, &w[t]);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    /* lrf_gather (nonbondene.f90:616-623): sum phi over workers, every worker gets the sum */
    if (sys->use_LRF && nthreads > 1) {
        int ig;
        size_t m;
        for (ig = 0; ig < sys->ncgp; ig++) {
            double *a = (double *)&w[0].st->lrf[ig];
            for (t = 1; t < nthreads; t++) {
                double *b = (double *)&w[t].st->lrf[ig];
                for (m = 3; m < QNB_LRF_STRIDE; m++) a[m] += b[m];
            }
        }
        for (t = 1; t < nthreads; t++) memcpy(w[t].st->lrf, w[0].st->lrf, sizeof(lrf_t) * (size_t)sys->ncgp);
    }
    for (t = 0; t < nthreads; t++) if (w[t].list_s > lmax) lmax = w[t].list_s;
    if (list_seconds) *list_seconds = lmax;
    t0 = now_s();
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker_force, &w[t]);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    /* gather_nonbond + serial sum on the master (potene.f90:200-222), inside the timed region */
    for (t = 1; t < nthreads; t++) {
        size_t m;
        for (m = 0; m < n3; m++) w[0].d[m] += w[t].d[m];
        for (k = 0; k < QNB_E_COUNT; k++) w[0].E[k] += w[t].E[k];
        for (k = 0; k < QNB_EQ_STRIDE * sys->nstates; k++) w[0].EQ[k] += w[t].EQ[k];
    }
    wall = now_s() - t0;
    if (d_out) memcpy(d_out, w[0].d, n3 * 8);
    if (E_out) memcpy(E_out, w[0].E, sizeof w[0].E);
    if (EQ_out) memcpy(EQ_out, w[0].EQ, sizeof(double) * QNB_EQ_STRIDE * (size_t)sys->nstates);
    for (t = 0; t < nthreads; t++) { qo_destroy(w[t].st); free(w[t].d); free(w[t].EQ); }
    free(w); free(th);
    return wall;
}


/* ------------------------------------------------------------------------------------------------
 * restrain_solvent (nonbondene.f90:6466-6543, solv_atom == 3 branch) and watpol (L6547-6746, SPC/TIP3P branch).
 * Literal restatement, same loop orders.
 */
int qo_solvent_restraints(const qnb_system *sys, const qnb_solvent_restraints *p, const double *theta_corr,
                          const double *x, int md, double *d, double E[2], double *shell_theta_sum, int32_t *shell_n) {
    const int nwat = sys->nwat, ns = sys->nat_solute, nsh = p->nwpolr_shell;
    const double pi = 3.14159265358979323846;
    /* ---- restrain_solvent L6486-6509: iw over the water charge groups, i = switch atom = O */
    for (int iw = 0; iw < nwat; iw++) {
        const int i = ns + 3 * iw;                       /* 0-based O */
        if (sys->excl[i]) continue;
        /* q_dist5(xwcent, x(i)): vec = x(i) - xwcent? -- math.f90: q_dist5(a,b)%vec = b - a */
        const double vx = x[3 * i] - p->xwcent[0], vy = x[3 * i + 1] - p->xwcent[1], vz = x[3 * i + 2] - p->xwcent[2];
        const double r2 = vx * vx + vy * vy + vz * vz;
        const double b = sqrt(r2);
        const double db = b - (p->rwat - p->shift);
        double erst, dv;
        if (db > 0.0) {
            erst = 0.5 * p->fk_wsphere * db * db - p->Dwmz;
            dv = p->fk_wsphere * db / b;
        } else if (b > 0.0) {
            const double fexp = exp(p->awmz * db);
            erst = p->Dwmz * (fexp * fexp - 2.0 * fexp);
            dv = -2.0 * p->Dwmz * p->awmz * (fexp - fexp * fexp) / b;
        } else { dv = 0.0; erst = 0.0; }
        d[3 * i] += vx * dv; d[3 * i + 1] += vy * dv; d[3 * i + 2] += vz * dv;
        E[0] += erst;
    }
    for (int is = 0; is < nsh; is++) { shell_theta_sum[is] = 0.0; shell_n[is] = 0; }
    if (!p->wpol_restr || nsh <= 0) return 0;
    /* ---- watpol */
    double *theta = (double *)calloc((size_t)(nwat > 0 ? nwat : 1), sizeof(double));
    double *tdum = (double *)calloc((size_t)(nwat > 0 ? nwat : 1), sizeof(double));
    int *list_sh = (int *)malloc(sizeof(int) * (size_t)(nwat > 0 ? nwat : 1) * (size_t)nsh);
    int *nsort = (int *)malloc(sizeof(int) * (size_t)(nwat > 0 ? nwat : 1) * (size_t)nsh);
    int *n_insh = (int *)calloc((size_t)nsh, sizeof(int));
    if (!theta || !tdum || !list_sh || !nsort || !n_insh) return 1;
    /* the sort happens "if (md)"; with md false the previous step's nsort would be reused -- potene only calls watpol
     * with in_md true (potene.f90:161-167), so that branch is not restated */
    (void)md;
    for (int iw = 0; iw < nwat; iw++) {                  /* L6565-6625 */
        theta[iw] = 0.0;
        const int i = ns + 3 * iw;
        if (sys->excl[i]) continue;
        double rmu[3], rcu[3];
        for (int c = 0; c < 3; c++) rmu[c] = (x[3 * (i + 1) + c] + x[3 * (i + 2) + c]) - x[3 * i + c] * 2.0;
        const double rm = sqrt(rmu[0] * rmu[0] + rmu[1] * rmu[1] + rmu[2] * rmu[2]);
        for (int c = 0; c < 3; c++) rmu[c] /= rm;
        for (int c = 0; c < 3; c++) rcu[c] = x[3 * i + c] - p->xwcent[c];
        const double rc = sqrt(rcu[0] * rcu[0] + rcu[1] * rcu[1] + rcu[2] * rcu[2]);
        for (int c = 0; c < 3; c++) rcu[c] /= rc;
        double scp = rmu[0] * rcu[0] + rmu[1] * rcu[1] + rmu[2] * rcu[2];
        if (scp > 1.0) scp = 1.0;
        if (scp < -1.0) scp = -1.0;
        theta[iw] = acos(scp);
        tdum[iw] = theta[iw];
        if (rc > p->rout[nsh - 1] - p->dr[nsh - 1]) {
            int is;
            for (is = nsh; is >= 2; is--)                /* do is = nwpolr_shell, 2, -1; exit if rc <= rout(is) */
                if (rc <= p->rout[is - 1]) break;
            /* after a completed Fortran loop is == 1 */
            list_sh[(size_t)(is - 1) * nwat + n_insh[is - 1]] = iw;
            n_insh[is - 1]++;
        }
    }
    for (int is = 0; is < nsh; is++) {                   /* L6627-6643: selection sort by theta, first minimum wins */
        int jmin = 0;
        for (int il = 0; il < n_insh[is]; il++) {
            double tmin = 2.0 * pi;
            for (int jl = 0; jl < n_insh[is]; jl++) {
                const int jw = list_sh[(size_t)is * nwat + jl];
                if (tdum[jw] < tmin) { jmin = jw; tmin = theta[jw]; }
            }
            nsort[(size_t)is * nwat + il] = jmin;
            tdum[jmin] = 99999.0;
        }
    }
    for (int is = 0; is < nsh; is++) {                   /* L6663-6742 */
        if (n_insh[is] == 0) continue;
        double avtdum = 0.0;
        for (int il = 1; il <= n_insh[is]; il++) {
            const int iw = nsort[(size_t)is * nwat + il - 1];
            const double arg = 1.0 + (1.0 - 2.0 * (double)il) / (double)n_insh[is];
            double theta0 = acos(arg);
            theta0 = theta0 - 3.0 * sin(theta0) * p->cstb[is] / 2.0;
            if (theta0 < 0.0) theta0 = 0.0;
            if (theta0 > pi) theta0 = pi;
            avtdum += theta[iw];
            const double dth = theta[iw] - theta0 + theta_corr[is];
            E[1] += 0.5 * p->fkwpol * dth * dth;
            const double dv = p->fkwpol * dth;
            const int i = ns + 3 * iw;
            double rmu[3], rcu[3];
            for (int c = 0; c < 3; c++) rmu[c] = (x[3 * (i + 1) + c] + x[3 * (i + 2) + c]) - x[3 * i + c] * 2.0;
            const double rm = sqrt(rmu[0] * rmu[0] + rmu[1] * rmu[1] + rmu[2] * rmu[2]);
            for (int c = 0; c < 3; c++) rmu[c] /= rm;
            for (int c = 0; c < 3; c++) rcu[c] = x[3 * i + c] - p->xwcent[c];
            const double rc = sqrt(rcu[0] * rcu[0] + rcu[1] * rcu[1] + rcu[2] * rcu[2]);
            for (int c = 0; c < 3; c++) rcu[c] /= rc;
            double scp = rmu[0] * rcu[0] + rmu[1] * rcu[1] + rmu[2] * rcu[2];
            if (scp > 1.0) scp = 1.0;
            if (scp < -1.0) scp = -1.0;
            double f0 = sin(acos(scp));
            if (fabs(f0) < 1.0e-10) f0 = 1.0e-10;   /* QREAL_EPS, sizes.f90:50 (double precision build) */
            f0 = -1.0 / f0;
            f0 = dv * f0;
            for (int c = 0; c < 3; c++) {
                const double f1 = ((rcu[c] - rmu[c] * scp) * (-2.0)) / rm;
                const double f3 = (rcu[c] - rmu[c] * scp) / rm;
                const double f2 = (rmu[c] - rcu[c] * scp) / rc;
                d[3 * i + c] += (f1 + f2) * f0;
                d[3 * (i + 1) + c] += f3 * f0;
                d[3 * (i + 2) + c] += f3 * f0;
            }
        }
        shell_theta_sum[is] = avtdum;
        shell_n[is] = n_insh[is];
    }
    free(theta); free(tdum); free(list_sh); free(nsort); free(n_insh);
    return 0;
}
