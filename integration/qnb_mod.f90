! qnb_mod.f90 -- ISO_C_BINDING interface to libqnb.so (include/qnb.h) for Qdyn6.
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is installed there); it is the
! binding a Q6 maintainer adds to src/ and lists in the makefile before nonbondene.f90.
! Every interface below mirrors one prototype of include/qnb.h.
! Coordinate and gradient arrays are TYPE(qr_vec) (three reals, sizes.f90:109-111) in Q6; they are passed as
! assumed-type, assumed-size dummies (type(*), dimension(*): Fortran 2018 / TS 29113, gfortran >= 4.9), which hands the
! C side the address of the first element without a type mismatch against real(c_double).
module QNB
use, intrinsic :: iso_c_binding
implicit none

integer(c_int), parameter :: QNB_ABI_VERSION = 1
integer(c_int), parameter :: QNB_IPC_BLOB = 128       ! bytes of a qnb_comm_ipc_export descriptor
integer(c_int), parameter :: QNB_FLAG_MD = 1, QNB_FLAG_QQ = 2, QNB_FLAG_NO_ENERGY = 4, QNB_FLAG_D_IS_ZERO = 8, &
                             QNB_FLAG_SOLVENT_RESTRAINTS = 16

! struct qnb_system (include/qnb.h) -- field order and types must match exactly
type, bind(c) :: qnb_system
    integer(c_int32_t) :: abi_version
    integer(c_int32_t) :: natom, nat_solute, nwat, solv_atom, ncgp, ncgp_solute, nqat, nstates, qswitch
    integer(c_int32_t) :: natyps, num_atyp, max_nbr_range, nexlong, n14long, nqlib, nqexpnb, nel_scale
    integer(c_int32_t) :: iuse_switch_atom, use_PBC, use_LRF, ivdw_rule, solvent_type, qvdw_flag
    integer(c_int32_t) :: qq_use_library_charges, ntors_gt_solute
    real(c_double)     :: el14_scale
    real(c_double)     :: xpcent(3)
    real(c_double)     :: rexcl_o
    type(c_ptr) :: cgp, cgpatom, excl, iqatom, iqseq, iac, crg, iaclib, ljcod, listex, list14, listexlong, list14long
    type(c_ptr) :: qcrg, qiac, qavdw, qbvdw, sc_lookup, iqexpnb, jqexpnb, el_scale_iq, el_scale_jq, el_scale, qconn
    integer(c_int32_t) :: pp_start, pp_end, pw_start, pw_end, qp_start, qp_end, ww_start, ww_end, qw_start, qw_end
    integer(c_int32_t) :: natom_start, natom_end, is_master
end type qnb_system

interface
    function qnb_last_error() bind(c, name='qnb_last_error') result(msg)
        import :: c_ptr
        type(c_ptr) :: msg
    end function
    function qnb_device_count() bind(c, name='qnb_device_count') result(n)
        import :: c_int
        integer(c_int) :: n                               ! usable CUDA devices on this node (0: every call fails)
    end function
    function qnb_init(sys, device, handle) bind(c, name='qnb_init') result(rc)
        import :: c_int, c_ptr, qnb_system
        type(qnb_system), intent(in) :: sys
        integer(c_int), value :: device
        type(c_ptr), intent(out) :: handle
        integer(c_int) :: rc
    end function
    function qnb_update_box(handle, boxlength, inv_boxl) bind(c, name='qnb_update_box') result(rc)
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: handle
        real(c_double), intent(in) :: boxlength(3), inv_boxl(3)
        integer(c_int) :: rc
    end function
    ! MC_volume (md.f90:1976): "old_nbww = nbww ... old_lrf = lrf" / the restore of a rejected move
    function qnb_save_lists(handle) bind(c, name='qnb_save_lists') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        integer(c_int) :: rc
    end function
    function qnb_restore_lists(handle) bind(c, name='qnb_restore_lists') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        integer(c_int) :: rc
    end function
    ! restrain_solvent / watpol (nonbondene.f90:6466-6746) inside the device step: type(qnb_solvent_restraints) mirrors
    ! include/qnb.h field for field (xwcent(3), rwat, fk_wsphere, shift, Dwmz, awmz, fkwpol, wpol_restr, nwpolr_shell,
    ! rout(8), dr(8), cstb(8)); the wrapper of watpol keeps wshell%avtheta / avn_insh / theta_corr on the host
    function qnb_set_solvent_restraints(handle, p) bind(c, name='qnb_set_solvent_restraints') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        type(c_ptr), value :: p                           ! c_loc of a type(qnb_solvent_restraints)
        integer(c_int) :: rc
    end function
    function qnb_set_theta_corr(handle, theta_corr) bind(c, name='qnb_set_theta_corr') result(rc)
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: handle
        real(c_double), intent(in) :: theta_corr(*)       ! wshell(1:nwpolr_shell)%theta_corr
        integer(c_int) :: rc
    end function
    function qnb_last_restraints(handle, E, shell_theta_sum, shell_n) bind(c, name='qnb_last_restraints') result(rc)
        import :: c_int, c_ptr, c_double, c_int32_t
        type(c_ptr), value :: handle
        real(c_double), intent(out) :: E(2)               ! solvent_radial, water_pol
        real(c_double), intent(out) :: shell_theta_sum(*) ! avtdum per shell
        integer(c_int32_t), intent(out) :: shell_n(*)     ! wshell(:)%n_insh
        integer(c_int) :: rc
    end function
    ! qcp_run (qcp.f90:319-372, 478-525): all beads of one sampling step in one call
    function qnb_qcp_beads(handle, x_save, natq, atoms, nbeads, coord, lambda, EQ_out) bind(c, name='qnb_qcp_beads') result(rc)
        import :: c_int, c_ptr, c_double, c_int32_t
        type(c_ptr), value :: handle
        type(*), dimension(*), intent(in) :: x_save       ! TYPE(qr_vec) x_save(natom) == 3*natom doubles
        integer(c_int), value :: natq, nbeads
        integer(c_int32_t), intent(in) :: atoms(*)        ! iqseq(qcp_atom(1:qcp_atnum))
        real(c_double), intent(in) :: coord(*)            ! qcp_coord(j,i) copied bead-major: (3, natq, nbeads)
        real(c_double), intent(in) :: lambda(*)
        real(c_double), intent(out) :: EQ_out(*)          ! (6, nstates, nbeads): qq, qp, qw x (el, vdw)
        integer(c_int) :: rc
    end function
    ! solvent SHAKE (bondene.f90:1069): constraints of the solvent molecules once, then shake(xx, x) per step
    function qnb_set_constraints(handle, nmol, mol_first, ij, dist2, winv) bind(c, name='qnb_set_constraints') result(rc)
        import :: c_int, c_ptr, c_double, c_int32_t
        type(c_ptr), value :: handle
        integer(c_int), value :: nmol
        integer(c_int32_t), intent(in) :: mol_first(*)    ! (nmol+1) first constraint of every constrained molecule, 0-based
        integer(c_int32_t), intent(in) :: ij(*)           ! const_mol(m)%bond%bond(ic)%i, %j interleaved, 1-based atoms
        real(c_double), intent(in) :: dist2(*)
        real(c_double), intent(in) :: winv(*)             ! winv(natom)
        integer(c_int) :: rc
    end function
    function qnb_shake(handle, xx, x, iterations) bind(c, name='qnb_shake') result(rc)
        import :: c_int, c_ptr, c_double, c_int64_t
        type(c_ptr), value :: handle
        type(c_ptr), value :: xx                          ! c_loc(xx) or c_null_ptr: the coordinates of this step's qnb_nonbond
        type(*), dimension(*), intent(inout) :: x         ! TYPE(qr_vec) x(natom)
        integer(c_int64_t), intent(out) :: iterations     ! sweeps summed over molecules (shake = iterations/nmol)
        integer(c_int) :: rc
    end function
    function qnb_build_lists(handle, x, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF, counts) &
            bind(c, name='qnb_build_lists') result(rc)
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: handle
        type(*), dimension(*), intent(in) :: x            ! TYPE(qr_vec) x(natom) == 3*natom doubles
        real(c_double), value :: Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF
        type(c_ptr), value :: counts                      ! c_null_ptr or int64(8)
        integer(c_int) :: rc
    end function
    function qnb_nonbond(handle, x, lambda, flags, d, E_out, EQ_out) bind(c, name='qnb_nonbond') result(rc)
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: handle
        type(*), dimension(*), intent(in) :: x            ! TYPE(qr_vec) x(natom)
        real(c_double), intent(in) :: lambda(*)
        integer(c_int), value :: flags
        type(*), dimension(*), intent(inout) :: d         ! TYPE(qr_vec) d(natom): added to
        real(c_double), intent(out) :: E_out(7), EQ_out(*) ! 6*nstates
        integer(c_int) :: rc
    end function
    ! page-lock x and d (module arrays of md.f90) once: no staging copy of x, the device adds the gradient into d
    function qnb_register_host_buffers(handle, x, d) bind(c, name='qnb_register_host_buffers') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        type(*), dimension(*), intent(in) :: x            ! TYPE(qr_vec) x(natom)
        type(*), dimension(*), intent(inout) :: d         ! TYPE(qr_vec) d(natom)
        integer(c_int) :: rc
    end function
    function qnb_release_host_buffers(handle) bind(c, name='qnb_release_host_buffers') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        integer(c_int) :: rc
    end function
    ! several independent systems of one process (lambda windows of a FEP farm, EVB frames) advanced together on one
    ! GPU: arrays of n handles / c_loc addresses, one entry per system; results are those of n single calls
    function qnb_build_lists_batch(n, handles, x, Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF, counts) &
            bind(c, name='qnb_build_lists_batch') result(rc)
        import :: c_int, c_ptr, c_double
        integer(c_int), value :: n
        type(c_ptr), intent(in) :: handles(*), x(*)       ! x(k) = c_loc of system k's coordinates
        real(c_double), value :: Rq, Rcq2, RcLRF2, Rcpp2, Rcpw2, Rcww2, RcLRF
        type(c_ptr), value :: counts                      ! c_null_ptr or int64(8, n)
        integer(c_int) :: rc
    end function
    function qnb_nonbond_batch(n, handles, x, lambda, flags, d, E_out, EQ_out) bind(c, name='qnb_nonbond_batch') result(rc)
        import :: c_int, c_ptr
        integer(c_int), value :: n
        type(c_ptr), intent(in) :: handles(*), x(*), lambda(*), d(*), E_out(*), EQ_out(*)   ! c_loc addresses per system
        integer(c_int), value :: flags
        integer(c_int) :: rc
    end function
    ! list sizes / explicit lists as the reference holds them (-DDUMP logging, debugging, nbmonitor-style analysis)
    function qnb_list_count(handle, which, state, n) bind(c, name='qnb_list_count') result(rc)
        import :: c_int, c_ptr, c_int64_t
        type(c_ptr), value :: handle
        integer(c_int), value :: which, state             ! QNB_LIST_PP..QNB_LIST_QQP (0..6); state 1-based for the Q lists
        integer(c_int64_t), intent(out) :: n
        integer(c_int) :: rc
    end function
    function qnb_export_list(handle, which, state, ij, params, capacity) bind(c, name='qnb_export_list') result(rc)
        import :: c_int, c_ptr, c_int32_t, c_int64_t
        type(c_ptr), value :: handle
        integer(c_int), value :: which, state
        integer(c_int32_t), intent(out) :: ij(*)          ! (2, n): i, j as in NB_TYPE / NBQP_TYPE / NBQ_TYPE, 1-based
        type(c_ptr), value :: params                      ! c_loc of real(c_double) (4, n): vdWA, vdWB, elec, score; or c_null_ptr
        integer(c_int64_t), value :: capacity
        integer(c_int) :: rc
    end function
    function qnb_export_lrf(handle, lrf) bind(c, name='qnb_export_lrf') result(rc)
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: handle
        real(c_double), intent(out) :: lrf(*)             ! 43*ncgp, LRF_TYPE order
        integer(c_int) :: rc
    end function
    function qnb_comm_unique_id(id128) bind(c, name='qnb_comm_unique_id') result(rc)
        import :: c_int, c_char
        character(kind=c_char), intent(out) :: id128(128)
        integer(c_int) :: rc
    end function
    function qnb_comm_init(handle, rank, nranks, id128) bind(c, name='qnb_comm_init') result(rc)
        import :: c_int, c_ptr, c_char
        type(c_ptr), value :: handle
        integer(c_int), value :: rank, nranks
        character(kind=c_char), intent(in) :: id128(128)
        integer(c_int) :: rc
    end function
    ! the same sums over peer memory (NVLink / NVSwitch) instead of NCCL: every rank exports a 128-byte descriptor, the
    ! host gathers them with MPI_Allgather and every rank attaches all of them (qnb_comm_init is then not needed)
    function qnb_comm_ipc_export(handle, blob) bind(c, name='qnb_comm_ipc_export') result(rc)
        import :: c_int, c_ptr, c_char
        type(c_ptr), value :: handle
        character(kind=c_char), intent(out) :: blob(128)
        integer(c_int) :: rc
    end function
    function qnb_comm_status(handle) bind(c, name='qnb_comm_status') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        integer(c_int) :: rc
    end function
    function qnb_comm_ipc_attach(handle, rank, nranks, blobs) bind(c, name='qnb_comm_ipc_attach') result(rc)
        import :: c_int, c_ptr, c_char
        type(c_ptr), value :: handle
        integer(c_int), value :: rank, nranks
        character(kind=c_char), intent(in) :: blobs(*)    ! 128 * nranks, by rank
        integer(c_int) :: rc
    end function
    function qnb_finalize(handle) bind(c, name='qnb_finalize') result(rc)
        import :: c_int, c_ptr
        type(c_ptr), value :: handle
        integer(c_int) :: rc
    end function
end interface

type(c_ptr), save :: qnb_handle = c_null_ptr
real(c_double), allocatable, save :: qnb_EQ(:)     ! 6*nstates of the last qnb_nonbond (qq terms are added later)

end module QNB
