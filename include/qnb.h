/*
 * qnb.h -- C ABI of the B200-native nonbonded engine for Qdyn6 (qusers/Q6).
 *
 * The reference has no FFI: the nonbonded path is a set of Fortran module
 * procedures that talk through `use`-associated globals.  This header is what a
 * thin ISO_C_BINDING shim binds so that the unchanged call sites
 *     md.f90:1691        call make_pair_lists(Rcq,Rcq**2,RcLRF**2,Rcpp**2,Rcpw**2,Rcww**2)
 *     md.f90:1742        call pot_energy(E,EQ,.true.)
 *       potene.f90:129     call pot_energy_nonbonds(E_loc,EQ_loc,md)
 *       potene.f90:176-177 call nonbond_qq / nonbond_qqp
 * run on the GPU.  The Fortran-side stub is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, host pointers only, no CUDA or torch types in any signature;
 *   - every array is laid out exactly as the Fortran host holds it: column-major,
 *     indices stored INSIDE the data are 1-based, default LOGICAL = int32 (0/!=0),
 *     TYPE(qr_vec) = three contiguous doubles (sizes.f90:109-111);
 *   - all entry points return 0 on success; non-zero -> qnb_last_error() holds the
 *     text and the Fortran wrapper calls die() (qalloc.f90:609);
 *   - one caller thread per handle, one handle per GPU; calls are synchronous at
 *     return (the host needs d immediately for SHAKE / the integrator);
 *   - there is NO CPU fallback: if no CUDA device is usable every call fails.
 */
#ifndef QNB_H
#define QNB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QNB_ABI_VERSION 1

/* ivdw_rule (topo.f90:853) */
#define QNB_VDW_GEOMETRIC 1
#define QNB_VDW_ARITHMETIC 2
/* solvent_type (topo.f90:1041) */
#define QNB_SOLVENT_SPC 0
#define QNB_SOLVENT_ALLATOM 1
#define QNB_SOLVENT_GENERAL 2

/* flags of qnb_nonbond */
#define QNB_FLAG_MD 1 /* pot_energy(...,in_md=.true.): pp/pw/ww/LRF as well as Q terms (potene.f90:333-376) */
#define QNB_FLAG_QQ 2 /* also evaluate the static nbqq/nbqqp lists (master only, potene.f90:176-177) */
/* Extension (off by default; the reference accumulates E on every call): the caller does not need the pp/pw/ww
 * energies of this step -- md.f90 reads E only when mod(istep, iout_cycle / iene_cycle / itemp_cycle) == 0 -- so the
 * row kernels skip their FP64 energy code.  E_out[pp,pw,ww] return 0; gradient, LRF and every Q term are unchanged. */
#define QNB_FLAG_NO_ENERGY 4
/* d is zero on entry -- pot_energy clears it right before the nonbonded terms (d(:) = zero, potene.f90:109; they are the
 * first contribution) -- so the gradient may be written instead of added: with a registered d
 * (qnb_register_host_buffers) the result is copied straight into it, no read of d and no host loop. */
#define QNB_FLAG_D_IS_ZERO 8

/* which list for qnb_list_count / qnb_export_list */
#define QNB_LIST_PP 0  /* nbpp  (globals.f90:415) */
#define QNB_LIST_PW 1  /* nbpw  */
#define QNB_LIST_WW 2  /* nbww  */
#define QNB_LIST_QP 3  /* nbqp  (i = Q-atom number, j = topology atom) */
#define QNB_LIST_QW 4  /* nbqw  */
#define QNB_LIST_QQ 5  /* nbqq  (static, per state; i = iq, j = jq) */
#define QNB_LIST_QQP 6 /* nbqqp (static, per state; i = iq, j = topology atom) */

/* number of doubles per charge group in qnb_export_lrf: LRF_TYPE, globals.f90:503-509 */
#define QNB_LRF_STRIDE 43

/* layout of E_out in qnb_nonbond: NB_ENERGIES{el,vdw} of pp, pw, ww, then LRF (nrgy.f90:37-40,48-54) */
#define QNB_E_PP_EL 0
#define QNB_E_PP_VDW 1
#define QNB_E_PW_EL 2
#define QNB_E_PW_VDW 3
#define QNB_E_WW_EL 4
#define QNB_E_WW_VDW 5
#define QNB_E_LRF 6
#define QNB_E_COUNT 7
/* layout of EQ_out: per state 6 doubles: qq.el qq.vdw qp.el qp.vdw qw.el qw.vdw */
#define QNB_EQ_STRIDE 6

/*
 * Static description of the system, taken after precompute_interactions
 * (qdyn.f90:153).  Everything is read during qnb_init and never retained.
 */
typedef struct qnb_system {
    int32_t abi_version; /* QNB_ABI_VERSION */

    /* ---- sizes (topo.f90 / qatom.f90 globals) ---- */
    int32_t natom;       /* nat_pro */
    int32_t nat_solute;
    int32_t nwat;        /* (natom-nat_solute)/solv_atom, simprep.f90:4545 */
    int32_t solv_atom;
    int32_t ncgp;
    int32_t ncgp_solute;
    int32_t nqat;
    int32_t nstates;
    int32_t qswitch;     /* FEP [PBC] switching_atom (1-based topology atom), 0 if unused */
    int32_t natyps;      /* rows of iaclib */
    int32_t num_atyp;    /* maxval(iac): dimension of ljcod, simprep.f90:4553 */
    int32_t max_nbr_range; /* 25, topo.f90:37 */
    int32_t nexlong;
    int32_t n14long;
    int32_t nqlib;       /* rows of qavdw/qbvdw */
    int32_t nqexpnb;     /* [soft_pairs] */
    int32_t nel_scale;   /* [el_scale] */

    /* ---- switches ---- */
    int32_t iuse_switch_atom; /* topo.f90:815 */
    int32_t use_PBC;
    int32_t use_LRF;
    int32_t ivdw_rule;
    int32_t solvent_type;
    int32_t qvdw_flag;               /* qatom.f90:862 */
    int32_t qq_use_library_charges;  /* qatom.f90:382 */
    int32_t ntors_gt_solute;         /* ntors>ntors_solute: solvent has internal nonbonded (unsupported -> error) */

    double el14_scale;
    double xpcent[3]; /* sphere centre (topo.f90:1089) */
    double rexcl_o;   /* exclusion radius (topo.f90:1071) */

    /* ---- topology arrays ---- */
    const int32_t *cgp;      /* [3*ncgp]  CGP_TYPE{iswitch,first,last}, topo.f90:73-77 */
    const int32_t *cgpatom;  /* [natom] */
    const int32_t *excl;     /* [natom] logical */
    const int32_t *iqatom;   /* [natom] 0 or Q-atom number */
    const int32_t *iqseq;    /* [nqat] topology atom of each Q-atom */
    const int32_t *iac;      /* [natom] */
    const double *crg;       /* [natom] already scaled by sqrt(coulomb_constant), simprep.f90:3714 */
    const double *iaclib;    /* [7*natyps] IAC_TYPE{mass,avdw(3),bvdw(3)}, topo.f90:79-83;
                                bvdw already sqrt'ed for the arithmetic rule (simprep.f90:4567-4571) */
    const int32_t *ljcod;    /* [num_atyp*num_atyp] */
    const int32_t *listex;   /* [max_nbr_range*nat_solute] logical, (k,i) -> k-1+(i-1)*max_nbr_range */
    const int32_t *list14;   /* same shape */
    const int32_t *listexlong; /* [2*nexlong] */
    const int32_t *list14long; /* [2*n14long] */

    /* ---- Q-atom / FEP tables (qatom.f90 header) ---- */
    const double *qcrg;      /* [nqat*nstates] (iq,state), scaled like crg */
    const int32_t *qiac;     /* [nqat*nstates] */
    const double *qavdw;     /* [nqlib*3] (type,code) */
    const double *qbvdw;     /* [nqlib*3] */
    const double *sc_lookup; /* [nqat*(natyps+nqat)*nstates], qatom.f90:1867 */
    const int32_t *iqexpnb;  /* [nqexpnb] */
    const int32_t *jqexpnb;  /* [nqexpnb] */
    const int32_t *el_scale_iq; /* [nel_scale] qq_el_scale%iqat */
    const int32_t *el_scale_jq; /* [nel_scale] qq_el_scale%jqat */
    const double *el_scale;  /* [nel_scale*nstates] (entry,state) column-major */
    const int32_t *qconn;    /* [nstates*nat_solute*nqat] (state,atom,iq), make_qconn nonbondene.f90:3087 */

    /* ---- work assignment: calculation_assignment%{pp,pw,qp,ww,qw}%{start,end}
     *      (mpiglob.f90:33-43); pp/pw/qp over 1..ncgp_solute, ww/qw over 1..nwat ---- */
    int32_t pp_start, pp_end;
    int32_t pw_start, pw_end;
    int32_t qp_start, qp_end;
    int32_t ww_start, ww_end;
    int32_t qw_start, qw_end;
    int32_t natom_start, natom_end; /* calculation_assignment%natom: lrf_taylor range (nonbondene.f90:519) */
    int32_t is_master;              /* nodeid==0: owner of the static nbqq/nbqqp lists (simprep.f90:3262) */
} qnb_system;

typedef struct qnb_handle qnb_handle;

/* Last error text of the calling thread's most recent failing call. */
const char *qnb_last_error(void);

/* Number of usable CUDA devices (0 => every other call fails; no CPU path). */
int qnb_device_count(void);

/*
 * Upload the static tables to GPU `device` and build every per-pair parameter
 * table the reference builds in precompute_interactions (simprep.f90:2860-3589)
 * plus the static nbqq/nbqqp lists (nbqqlist, nonbondene.f90:3231).
 * Call once after precompute_interactions (qdyn.f90:153).
 */
int qnb_init(const qnb_system *sys, int device, qnb_handle **out);

/* md.f90 MC_volume / put_back_in_box: new box (boxlength, inv_boxl). Ignored unless use_PBC. */
int qnb_update_box(qnb_handle *h, const double boxlength[3], const double inv_boxl[3]);

/*
 * MC_volume (md.f90:1976-2284) keeps the pair lists, their counts and the LRF moments of the current box while it
 * tries a scaled one (old_nbww = nbww ... old_lrf(:) = lrf(:), md.f90:2022-2058) and puts them back when the move is
 * rejected (md.f90:2214-2256).  qnb_save_lists remembers the state of the last qnb_build_lists (coordinates the
 * lists were built from, cut-offs, box, LRF moments); qnb_restore_lists re-creates the device lists from it
 * (the lists are a pure function of that state) and restores the saved box and LRF moments bit for bit.
 */
int qnb_save_lists(qnb_handle *h);
int qnb_restore_lists(qnb_handle *h);

/*
 * Spherical-boundary solvent restraints, the per-step host neighbours of the nonbonded call inside pot_energy
 * (potene.f90:161-167): restrain_solvent (nonbondene.f90:6466-6543) and watpol
 * (nonbondene.f90:6547-6746) for three-site solvents.  The parameters are what the host prepared from the input
 * and in wat_shells (simprep.f90:4707-4891); wshell(:)%theta_corr stays host state (it changes every itdis steps,
 * nonbondene.f90:6640-6652) and is pushed with qnb_set_theta_corr.  With QNB_FLAG_SOLVENT_RESTRAINTS (and
 * QNB_FLAG_MD, sphere only) qnb_nonbond adds both terms to d inside the same device step -- no extra copies.
 */
#define QNB_MAX_SHELLS 8
typedef struct qnb_solvent_restraints {
    double xwcent[3];     /* solvent sphere centre */
    double rwat;          /* solvent sphere radius */
    double fk_wsphere;    /* radial force constant */
    double shift;         /* q_sqrt(Boltz*Tfree/fk_wsphere), zero when fk_wsphere == 0 (nonbondene.f90:6479-6483) */
    double Dwmz, awmz;    /* surface well */
    double fkwpol;        /* polarisation force constant */
    int32_t wpol_restr;   /* polarisation restraint switched on (potene.f90:167) */
    int32_t nwpolr_shell; /* 0..QNB_MAX_SHELLS */
    double rout[QNB_MAX_SHELLS], dr[QNB_MAX_SHELLS], cstb[QNB_MAX_SHELLS]; /* SHELL_TYPE, globals.f90:165 */
} qnb_solvent_restraints;
#define QNB_FLAG_SOLVENT_RESTRAINTS 16
int qnb_set_solvent_restraints(qnb_handle *h, const qnb_solvent_restraints *p);
int qnb_set_theta_corr(qnb_handle *h, const double *theta_corr /* [nwpolr_shell] */);
/* Results of the last qnb_nonbond that carried QNB_FLAG_SOLVENT_RESTRAINTS: E[0] = E%restraint%solvent_radial,
 * E[1] = E%restraint%water_pol; per shell the sum of theta over its molecules (avtdum, nonbondene.f90:6668) and
 * n_insh, from which the host keeps wshell%avtheta / avn_insh (L6738-6741). */
int qnb_last_restraints(qnb_handle *h, double E[2], double *shell_theta_sum, int32_t *shell_n);

/*
 * SHAKE of solvent-sized molecules: shake(xx, x) of bondene.f90:1069-1150 over the constraints
 * init_constraints (simprep.f90:2167-2345) builds -- with the default input the three constraints of every water.
 * After the nonbonded path is offloaded this is the largest per-step loop left on the host (md.f90: leap-frog update,
 * then `niter = constraint(const_method, xx, x)`).
 *   nmol, mol_first[nmol+1]  constrained molecules: constraints mol_first[m] .. mol_first[m+1]-1 (0-based, CSR) belong
 *                            to molecule m, in the order the reference sweeps them; at most 32 per molecule (solute
 *                            SHAKE -- one molecule with thousands of constraints -- stays on the host); molecules
 *                            must not share atoms
 *   ij[2*n], dist2[n]        const_mol%bond%bond(:)%i, %j (1-based atoms) and %dist2
 *   winv[natom]              inverse masses (simprep.f90:3675)
 * qnb_shake corrects x[3*natom] in place.  xx = reference coordinates; NULL = the coordinates already resident on the
 * device from this step's qnb_nonbond (the leap-frog case: xx is the x the forces were evaluated at), which saves one
 * upload.  iterations (may be NULL) receives the sweeps summed over molecules (the reference returns that sum / nmol).
 * A molecule that does not converge in CONST_MAX_ITER sweeps fails the call with "shake failure" (bondene.f90:1143).
 * Same order of operations as the reference in explicitly rounded FP64: results are bit-identical to a CPU SHAKE.
 */
int qnb_set_constraints(qnb_handle *h, int nmol, const int32_t *mol_first, const int32_t *ij, const double *dist2,
                        const double *winv);
int qnb_shake(qnb_handle *h, const double *xx, double *x, int64_t *iterations);

/*
 * qcp_run (qcp.f90:319-372, 478-525): for every bead i the coordinates of the path-integral atoms are set to
 * x(iqseq(qcp_atom(j))) = x_save(..) + qcp_coord(j,i) and pot_energy(qcp_E,qcp_EQ,.false.) is called; only the
 * per-state Q energies are used.  This entry evaluates the nonbonded part of all beads in one call (one upload of
 * x_save, the beads' displacements applied on the device, one download of the energies).
 *   atoms[natq]           1-based atom numbers of the displaced atoms (iqseq(qcp_atom(j)))
 *   coord[nbeads][natq][3] displacements (qcp_coord(j,i), bead-major)
 *   EQ_out[nbeads][6*nstates] as qnb_nonbond's EQ_out, per bead
 * Uses the current pair lists (QCP never rebuilds them, qcp.f90).  Forces are not produced.
 */
int qnb_qcp_beads(qnb_handle *h, const double *x_save, int natq, const int32_t *atoms, int nbeads, const double *coord,
                  const double *lambda, double *EQ_out);

/*
 * make_pair_lists (nonbondene.f90:749): rebuild nbpp/nbpw/nbww/nbqp/nbqw for the
 * handle's shard from coordinates x[3*natom] and accumulate the LRF moments.
 * Arguments are the reference's, in its order; RcLRF (unsquared) is the global
 * the box builders compare against -1 (nonbondene.f90:3066,4497).
 * counts_out (may be NULL): nbpp_pair, nbpw_pair, nbww_pair, nbqp_pair, nbqw_pair,
 * nbpp_cgp_pair, nbpw_cgp_pair, nbqp_cgp_pair (md.f90:1693-1695 logging).
 */
int qnb_build_lists(qnb_handle *h, const double *x, double Rq, double Rcq2, double RcLRF2,
                    double Rcpp2, double Rcpw2, double Rcww2, double RcLRF, int64_t counts_out[8]);

/*
 * pot_energy_nonbonds (potene.f90:320) + nonbond_qq/nonbond_qqp (potene.f90:176-177).
 * x[3*natom] in, lambda[nstates] in; d[3*natom] is ADDED to (the host zeroes it,
 * potene.f90:109, and adds bonded terms afterwards); E_out[QNB_E_COUNT] and
 * EQ_out[QNB_EQ_STRIDE*nstates] are overwritten with this call's sums.
 */
int qnb_nonbond(qnb_handle *h, const double *x, const double *lambda, int flags, double *d,
                double *E_out, double *EQ_out);
/*
 * Optional, once after qnb_init: page-lock (cudaHostRegister) the host's coordinate and gradient arrays -- Q6's x and d are
 * module arrays that live as long as the run (md.f90) -- so that qnb_nonbond / qnb_nonbond_batch calls made with exactly
 * these addresses upload x without a staging copy and let the device add the gradient into d (no host loop, no second
 * copy).  Calls with other addresses keep staging through the library's own pinned buffers.  Either pointer may be NULL.
 * The arrays stay registered until qnb_finalize; a host that frees or reallocates them earlier releases them first.
 * QNB_NO_HOST_REGISTER=1 in the environment turns both calls into no-ops.
 */
int qnb_register_host_buffers(qnb_handle *h, const double *x, double *d);
int qnb_release_host_buffers(qnb_handle *h);

/*
 * Several independent systems of one process -- FEP lambda windows (tests/exclude_tests/run_excl_test.sh:92-125 runs 51
 * of them as separate Qdyn6 jobs), EVB frames, replicas -- advanced together on one GPU: the same calls as
 * make_pair_lists / pot_energy_nonbonds for every system k = 0..n-1 (its own handle, coordinates, lambda, gradient and
 * energy arrays), issued so that the device works on all of them at once.  Results are those of n single calls.
 */
int qnb_build_lists_batch(int n, qnb_handle *const *h, const double *const *x, double Rq, double Rcq2, double RcLRF2,
                          double Rcpp2, double Rcpw2, double Rcww2, double RcLRF, int64_t *counts_out /* [n][8] or NULL */);
int qnb_nonbond_batch(int n, qnb_handle *const *h, const double *const *x, const double *const *lambda, int flags,
                      double *const *d, double *const *E_out, double *const *EQ_out);

/* Number of entries of a list as the reference would hold it (per state for Q lists). */
int qnb_list_count(qnb_handle *h, int which, int state /*1-based, Q lists*/, int64_t *n);
/*
 * Expand a device list into the reference's explicit entries (order unspecified:
 * compare as sorted sets).  ij receives 2*n int32 (i,j interleaved, 1-based, i/j as in
 * NB_TYPE/NBQP_TYPE/NBQ_TYPE); params (may be NULL) receives 4*n doubles
 * vdWA,vdWB,elec,score for `state`.
 */
int qnb_export_list(qnb_handle *h, int which, int state, int32_t *ij, double *params, int64_t capacity);

/* lrf(1:ncgp) as LRF_TYPE: cgp_cent(3), phi0, phi1(3), phi2(3x3), phi3(9x3) */
int qnb_export_lrf(qnb_handle *h, double *lrf);

/* ---- multi-GPU: i-range sharding with an all-reduce (gather_nonbond / lrf_gather) ---- */
/* 128-byte NCCL unique id; rank 0 creates it, the host broadcasts it (MPI_Bcast / torch.distributed). */
int qnb_comm_unique_id(void *id128);
int qnb_comm_init(qnb_handle *h, int rank, int nranks, const void *id128);

/*
 * The same sums over peer memory instead of NCCL, for ranks on one NVLink / NVSwitch node: every rank exports a
 * QNB_IPC_BLOB-byte descriptor of its device buffers, the host gathers them (MPI_Allgather / torch.distributed) and every
 * rank attaches all of them.  From then on the per-step sum of [d | E | EQ] and the per-build sum of the LRF moments are
 * one kernel per rank that reads and writes the peers' buffers directly (captured in the step's CUDA graph); qnb_comm_init
 * is not needed.  Fails (use qnb_comm_init) when the devices cannot access each other.
 */
#define QNB_IPC_BLOB 128
int qnb_comm_ipc_export(qnb_handle *h, void *blob /* QNB_IPC_BLOB bytes */);
int qnb_comm_ipc_attach(qnb_handle *h, int rank, int nranks, const void *blobs /* nranks * QNB_IPC_BLOB bytes, by rank */);
/* 0, or non-zero (qnb_last_error) when a rank missed a barrier of the peer-memory all-reduce (it gives up after ~4 s). */
int qnb_comm_status(qnb_handle *h);

/* ---- measurement hooks (bench.py); device-resident, no host copies ---- */
/* Re-run the last qnb_nonbond `steps` times on the coordinates already in HBM;
 * ms_out = CUDA-event time of the whole loop on the handle's stream. */
int qnb_bench_nonbond(qnb_handle *h, const double *lambda, int flags, int steps, int flush_l2, float *ms_out);
int qnb_bench_build_lists(qnb_handle *h, int reps, float *ms_out);
/* Parts of the builds timed by the last qnb_bench_build_lists, ms per build (CUDA events on the streams the kernels
 * run on): LRF accumulation (lrf_update, on its side stream), row scan count pass, row scan fill pass. */
int qnb_bench_last_build_timing(qnb_handle *h, float out[3]);
/* md_run's loop on device-resident coordinates: a list rebuild every nbcycle steps (md.f90:1661; the step counter runs
 * on across calls) plus one nonbonded evaluation per step; ms_out = CUDA-event time of the loop. */
int qnb_bench_md(qnb_handle *h, const double *lambda, int flags, int steps, int nbcycle, float *ms_out);
/* Sustained FMA throughput of the FP32 (which=0) / FP64 (which=1) pipe in TFLOP/s: the measured roofline
 * denominator of the force kernels (MEASURED_PEAKS.json holds only HBM and bf16 tensor figures). */
int qnb_bench_peak(int device, int which, float *tflops_out);
/* Per-kernel CUDA-event time (ms, averaged over reps) of the kernels of the last step:
 * names_out: '\n'-separated list; returns number of kernels. */
int qnb_bench_kernels(qnb_handle *h, const double *lambda, int flags, int reps, int flush_l2,
                      char *names_out, int names_cap, float *ms_out, int ms_cap);
/* Host seconds of the last qnb_nonbond_batch in the calling thread: issuing (staging + graph launches), waiting for the
 * device, adding the gradients out. */
int qnb_bench_last_batch_timing(double out[3]);
/* ms per all-reduce of [d | E | EQ] alone (sharded handles; the collective that replaces gather_nonbond). */
int qnb_bench_allreduce(qnb_handle *h, int reps, float *ms_out);
/* Phases of the last peer-memory all-reduce on this rank, us (%globaltimer): waiting for all ranks to arrive, summing and
 * delivering the own slice, waiting for all ranks to finish. */
int qnb_bench_allreduce_phases(qnb_handle *h, double out[3]);
/* Kernel launches issued by this handle since creation. */
int64_t qnb_launch_count(qnb_handle *h);
/* Host-side seconds of the last qnb_nonbond: staging x into pinned memory, issuing the step (graph launch),
 * waiting for the device (H2D + kernels + D2H), adding d / summing energies. */
int qnb_last_timing(qnb_handle *h, double out[4]);
/* Bytes copied host->device / device->host by the last qnb_nonbond call. */
int qnb_last_copy_bytes(qnb_handle *h, int64_t *h2d, int64_t *d2h);

int qnb_finalize(qnb_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* QNB_H */
