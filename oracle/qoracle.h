/*
 * qoracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU (FP64, single thread per state) restatement of the Qdyn6 nonbonded path of
 * qusers/Q6, written from the Fortran sources function by function (each C
 * function cites the file:line it follows).  It is the parity authority for the
 * CUDA library in q6_b200/csrc: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * never links or calls it.
 *
 * Pinning: the reference cannot be compiled in the build container (no Fortran
 * compiler), so the oracle is pinned on the reference's own golden values
 * (tests/test_oracle_golden.py): row 1 of tests/basic_tests/{SPH,PBC}_*_benchmark.en
 * and eval_test.sh:99-101 -- the step-0 energies QEL, QVdW (qp+qw) and EL, VdW
 * (pp+pw+ww) of the shipped sphere and periodic runs -- are reproduced to the
 * printed digit (3.12 139.43 -7.30 -413.17 / -31.30 228.64 380.33 -1990.55) at the
 * coordinates the reference evaluates them at, i.e. after its initial solvent
 * SHAKE (restated in oracle/pyoracle.py: shake), plus 56 402 water pairs inside
 * 10 A.  That pins the ww and qw list builders and energy sums (sphere and box).
 * Forces, the pp/pw/qp/qq terms, LRF moments and multi-state energies are held by
 * no reference test: "parity unpinned" there, beyond self-consistency checks
 * (finite differences, LRF convergence); DESIGN.md says so too.
 */
#ifndef QORACLE_H
#define QORACLE_H

#include "../include/qnb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qo_state qo_state;

/* precompute_interactions + make_nbqqlist/nbqqlist on a copy of the tables */
qo_state *qo_create(const qnb_system *sys);
void qo_destroy(qo_state *s);
const char *qo_last_error(void);

void qo_update_box(qo_state *s, const double boxlength[3], const double inv_boxl[3]);

/* make_pair_lists, nonbondene.f90:749 */
int qo_make_pair_lists(qo_state *s, const double *x, double Rq, double Rcq2, double RcLRF2, double Rcpp2,
                       double Rcpw2, double Rcww2, double RcLRF, int64_t counts_out[8]);

/* pot_energy_nonbonds (potene.f90:320) [+ nonbond_qq/nonbond_qqp when QNB_FLAG_QQ]; d += */
int qo_nonbond(qo_state *s, const double *x, const double *lambda, int flags, double *d, double *E_out,
               double *EQ_out);

int qo_list_count(qo_state *s, int which, int state, int64_t *n);
/* entries in the reference's own order */
int qo_export_list(qo_state *s, int which, int state, int32_t *ij, double *params, int64_t capacity);
int qo_export_lrf(qo_state *s, double *lrf);

/* restrain_solvent (nonbondene.f90:6466-6543) + watpol (L6547-6746) for solv_atom == 3; d +=, E[2] +=;
 * shell_theta_sum / shell_n as qnb_last_restraints.  md as watpol's argument. */
int qo_solvent_restraints(const qnb_system *sys, const qnb_solvent_restraints *p, const double *theta_corr,
                          const double *x, int md, double *d, double E[2], double *shell_theta_sum, int32_t *shell_n);

/* make_qconn / find_bonded (nonbondene.f90:3087-3187), for validating the host-side port.
 * bnd: [3*nbonds_solute] i,j,cod; qbnd_ij: [2*nqbond]; qbnd_cod: [nqbond*nstates] (bond,state);
 * exspec_ij: [2*nexspec]; exspec_flag: [nexspec*nstates]; out: qconn[nstates*nat_solute*nqat] */
void qo_make_qconn(int nstates, int nat_solute, int nqat, const int32_t *iqseq, const int32_t *iqatom,
                   int nbonds_solute, const int32_t *bnd, int nqbond, const int32_t *qbnd_ij,
                   const int32_t *qbnd_cod, int nexspec, const int32_t *exspec_ij, const int32_t *exspec_flag,
                   int32_t *qconn);

/*
 * Decomposed run (the reference's Qdyn6p: distribute_nonbonds, nonbondene.f90:80-505):
 * nthreads states each owning an i-range, private force arrays summed at the end.
 * Returns wall seconds of `steps` nonbond evaluations (lists built once before timing
 * when build_each==0, else rebuilt every step).
 */
double qo_time_decomposed(const qnb_system *sys, const double *x, const double *lambda, int flags,
                          const double cut[7] /*Rq,Rcq2,RcLRF2,Rcpp2,Rcpw2,Rcww2,RcLRF*/,
                          const double *box6 /*boxlength,inv_boxl or NULL*/, int nthreads, int steps,
                          int build_each, double *d_out, double *E_out, double *EQ_out, double *list_seconds);

#ifdef __cplusplus
}
#endif
#endif
