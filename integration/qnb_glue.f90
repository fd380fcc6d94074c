! qnb_glue.f90 -- the bodies that Q6's make_pair_lists / pot_energy_nonbonds / nonbond_qq / nonbond_qqp get when Qdyn6 is
! built with -DUSE_QNB (see integration/q6_qnb.patch, which adds the three-line #if branches that call into here).
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler).  It is written against the module globals of
! qusers/Q6 src/ (topo.f90, qatom.f90, globals.f90, mpiglob.f90, nrgy.f90) and the interfaces of qnb_mod.f90; the C++
! host q6_b200/host/qdyn_host.cpp (prep_sim) fills the same struct from the same tables and IS compiled and tested, so
! it is the working reference for what every field must hold.
!
! Everything qnb_system points to is read during qnb_init only.  c_loc needs TARGET, which Q6's allocatable globals do
! not carry, so the tables are copied once into the module-level target arrays below (a few MB, start-up only).
module QNB_GLUE
use, intrinsic :: iso_c_binding
use QNB
use SIZES
use TOPO
use QATOM
use GLOBALS
use QALLOC
use MPIGLOB
use NRGY
implicit none

integer(c_int32_t), allocatable, target, save, private :: g_cgp(:,:), g_cgpatom(:), g_excl(:), g_iqatom(:), g_iqseq(:), g_iac(:)
integer(c_int32_t), allocatable, target, save, private :: g_ljcod(:,:), g_listex(:,:), g_list14(:,:), g_exlong(:,:), g_14long(:,:)
integer(c_int32_t), allocatable, target, save, private :: g_qiac(:,:), g_iqexpnb(:), g_jqexpnb(:), g_els_i(:), g_els_j(:), g_qconn(:,:,:)
real(c_double), allocatable, target, save, private     :: g_crg(:), g_iaclib(:,:), g_qcrg(:,:), g_qavdw(:,:), g_qbvdw(:,:)
real(c_double), allocatable, target, save, private     :: g_sc(:,:,:), g_els(:,:)
! device = mod(nodeid, qnb_device_count()): ranks are dealt round-robin to the GPUs this node really has

contains

function qnb_message() result(msg)
  character(len=:), allocatable :: msg
  character(kind=c_char), pointer :: p(:)
  integer :: n
  call c_f_pointer(qnb_last_error(), p, (/ 1024 /))
  n = 0
  do while (n < 1024)
     if (p(n+1) == c_null_char) exit
     n = n + 1
  end do
  allocate(character(len=n) :: msg)
  msg = transfer(p(1:n), msg)
end function qnb_message

! --- once, after precompute_interactions (qdyn.f90:153)
subroutine qnb_setup
  type(qnb_system) :: s
  integer :: i, k, ngpu
  real(c_double) :: bl(3), ibl(3)

  ! The interfaces pass x, d and every real table as c_double: a Q6 built with -DQSINGLE / -DQUADRUPLE cannot use them.
  if (kind(1.0_prec) /= c_double) call die('USE_QNB needs the double-precision build of Q6 (-DQDOUBLE)')
  ! Qdyn6p (several MPI ranks, one GPU each): every rank evaluates its calculation_assignment share and the library sums
  ! d, E, EQ (per step) and the LRF moments (per list build) over the ranks - qnb_connect_ranks below.  The patch takes
  ! the reference's own exchange out of the USE_QNB build: make_pair_lists returns before lrf_gather
  ! (nonbondene.f90:831), gather_nonbond (L6760) returns at once and the master's sum over d_recv / E_recv
  ! (potene.f90:200-222) is compiled out, else the complete sums would be counted once per rank.
  ! MC_volume's restore of a rejected move runs on the master only (md.f90:2203): not wired for several ranks.
  if (numnodes > 1 .and. use_PBC .and. constant_pressure) &
     call die('USE_QNB: constant_pressure (MC_volume) with several MPI ranks is not hooked up in this glue')
  ngpu = qnb_device_count()
  if (ngpu < 1) call die('USE_QNB: no CUDA device: '//qnb_message())

  allocate(g_cgp(3,ncgp), g_cgpatom(size(cgpatom)), g_excl(natom), g_iqatom(natom), g_iac(natom), g_crg(natom))
  do i = 1, ncgp
     g_cgp(1,i) = cgp(i)%iswitch; g_cgp(2,i) = cgp(i)%first; g_cgp(3,i) = cgp(i)%last
  end do
  g_cgpatom = cgpatom
  g_excl = merge(1, 0, excl(1:natom))
  g_iqatom = iqatom(1:natom)
  g_iac = iac(1:natom)
  g_crg = crg(1:natom)                                ! already * sqrt(coulomb_constant) (simprep.f90:3714)
  allocate(g_iaclib(7,natyps))
  do i = 1, natyps
     g_iaclib(1,i) = iaclib(i)%mass; g_iaclib(2:4,i) = iaclib(i)%avdw(1:3); g_iaclib(5:7,i) = iaclib(i)%bvdw(1:3)
  end do
  allocate(g_ljcod(num_atyp,num_atyp)); g_ljcod = ljcod(1:num_atyp,1:num_atyp)
  allocate(g_listex(max_nbr_range,max(nat_solute,1)), g_list14(max_nbr_range,max(nat_solute,1)))
  g_listex = 0; g_list14 = 0
  if (nat_solute > 0) then
     g_listex(:,1:nat_solute) = merge(1, 0, listex(1:max_nbr_range,1:nat_solute))
     g_list14(:,1:nat_solute) = merge(1, 0, list14(1:max_nbr_range,1:nat_solute))
  end if
  allocate(g_exlong(2,max(nexlong,1)), g_14long(2,max(n14long,1)))
  if (nexlong > 0) g_exlong(:,1:nexlong) = listexlong(1:2,1:nexlong)
  if (n14long > 0) g_14long(:,1:n14long) = list14long(1:2,1:n14long)

  allocate(g_iqseq(max(nqat,1)), g_qcrg(max(nqat,1),max(nstates,1)), g_qiac(max(nqat,1),max(nstates,1)))
  allocate(g_qavdw(max(nqlib,1),3), g_qbvdw(max(nqlib,1),3), g_sc(max(nqat,1),natyps+max(nqat,1),max(nstates,1)))
  allocate(g_iqexpnb(max(nqexpnb,1)), g_jqexpnb(max(nqexpnb,1)))
  allocate(g_els_i(max(nel_scale,1)), g_els_j(max(nel_scale,1)), g_els(max(nel_scale,1),max(nstates,1)))
  allocate(g_qconn(max(nstates,1),max(nat_solute,1),max(nqat,1)))
  if (nqat > 0) then
     g_iqseq = iqseq(1:nqat)
     g_qcrg = qcrg(1:nqat,1:nstates)                  ! scaled like crg
     g_qiac = qiac(1:nqat,1:nstates)
     if (nqlib > 0) then
        g_qavdw = qavdw(1:nqlib,1:3); g_qbvdw = qbvdw(1:nqlib,1:3)
     end if
     g_sc = sc_lookup(1:nqat,1:natyps+nqat,1:nstates)
     if (nqexpnb > 0) then
        g_iqexpnb = iqexpnb(1:nqexpnb); g_jqexpnb = jqexpnb(1:nqexpnb)
     end if
     do k = 1, nel_scale
        g_els_i(k) = qq_el_scale(k)%iqat; g_els_j(k) = qq_el_scale(k)%jqat
        g_els(k,1:nstates) = qq_el_scale(k)%el_scale(1:nstates)
     end do
     g_qconn = qconn(1:nstates,1:nat_solute,1:nqat)
  end if

  s%abi_version = QNB_ABI_VERSION
  s%natom = natom; s%nat_solute = nat_solute; s%nwat = nwat; s%solv_atom = solv_atom
  s%ncgp = ncgp; s%ncgp_solute = ncgp_solute; s%nqat = nqat; s%nstates = nstates; s%qswitch = qswitch
  s%natyps = natyps; s%num_atyp = num_atyp; s%max_nbr_range = max_nbr_range
  s%nexlong = nexlong; s%n14long = n14long; s%nqlib = nqlib; s%nqexpnb = nqexpnb; s%nel_scale = nel_scale
  s%iuse_switch_atom = iuse_switch_atom; s%use_PBC = merge(1, 0, use_PBC); s%use_LRF = merge(1, 0, use_LRF)
  s%ivdw_rule = ivdw_rule; s%solvent_type = solvent_type
  s%qvdw_flag = merge(1, 0, qvdw_flag); s%qq_use_library_charges = merge(1, 0, qq_use_library_charges)
  s%ntors_gt_solute = merge(1, 0, ntors > ntors_solute)
  s%el14_scale = el14_scale; s%rexcl_o = rexcl_o
  s%xpcent = (/ xpcent%x, xpcent%y, xpcent%z /)
  s%cgp = c_loc(g_cgp); s%cgpatom = c_loc(g_cgpatom); s%excl = c_loc(g_excl); s%iqatom = c_loc(g_iqatom)
  s%iqseq = c_loc(g_iqseq); s%iac = c_loc(g_iac); s%crg = c_loc(g_crg); s%iaclib = c_loc(g_iaclib)
  s%ljcod = c_loc(g_ljcod); s%listex = c_loc(g_listex); s%list14 = c_loc(g_list14)
  s%listexlong = c_loc(g_exlong); s%list14long = c_loc(g_14long)
  s%qcrg = c_loc(g_qcrg); s%qiac = c_loc(g_qiac); s%qavdw = c_loc(g_qavdw); s%qbvdw = c_loc(g_qbvdw)
  s%sc_lookup = c_loc(g_sc); s%iqexpnb = c_loc(g_iqexpnb); s%jqexpnb = c_loc(g_jqexpnb)
  s%el_scale_iq = c_loc(g_els_i); s%el_scale_jq = c_loc(g_els_j); s%el_scale = c_loc(g_els); s%qconn = c_loc(g_qconn)
  s%pp_start = calculation_assignment%pp%start; s%pp_end = calculation_assignment%pp%end
  s%pw_start = calculation_assignment%pw%start; s%pw_end = calculation_assignment%pw%end
  s%qp_start = calculation_assignment%qp%start; s%qp_end = calculation_assignment%qp%end
  s%ww_start = calculation_assignment%ww%start; s%ww_end = calculation_assignment%ww%end
  s%qw_start = calculation_assignment%qw%start; s%qw_end = calculation_assignment%qw%end
  s%natom_start = calculation_assignment%natom%start; s%natom_end = calculation_assignment%natom%end
  s%is_master = merge(1, 0, nodeid == 0)

  if (qnb_init(s, mod(nodeid, ngpu), qnb_handle) /= 0) call die('qnb_init: '//qnb_message())
  if (numnodes > 1) call qnb_connect_ranks
  ! x and d are module arrays allocated once (md.f90): page-locked, so that every step uploads x without a staging copy
  ! and the device adds the gradient into d; a failure only means the staged path is used
  if (qnb_register_host_buffers(qnb_handle, x, d) /= 0) write(*,'(a)') 'qnb: x / d not page-locked: '//qnb_message()
  if (use_PBC) then
     bl = (/ boxlength%x, boxlength%y, boxlength%z /); ibl = (/ inv_boxl%x, inv_boxl%y, inv_boxl%z /)
     if (qnb_update_box(qnb_handle, bl, ibl) /= 0) call die('qnb_update_box: '//qnb_message())
  end if
  allocate(qnb_EQ(6*max(nstates,1)))
  qnb_EQ = 0
  ! the start-up copies are no longer needed: qnb_init does not retain host pointers
  deallocate(g_cgp, g_cgpatom, g_excl, g_iqatom, g_iqseq, g_iac, g_ljcod, g_listex, g_list14, g_exlong, g_14long, g_qiac, &
             g_iqexpnb, g_jqexpnb, g_els_i, g_els_j, g_qconn, g_crg, g_iaclib, g_qcrg, g_qavdw, g_qbvdw, g_sc, g_els)
end subroutine qnb_setup

! --- Qdyn6p: connect the GPUs of the ranks (once, from qnb_setup).  First choice is the all-reduce kernel over peer memory
! (NVLink / NVSwitch, ranks of one node): every rank exports a CUDA IPC descriptor of its arena, MPI_Allgather hands all of
! them to everybody, every rank maps its peers.  If any rank cannot (no peer access, ranks on several nodes) ALL ranks go
! to NCCL: rank 0 draws the unique id, MPI_Bcast, qnb_comm_init (which also drops a peer mapping that did succeed here).
subroutine qnb_connect_ranks
#if defined (USE_MPI)
  character(kind=c_char), target :: blob(QNB_IPC_BLOB), id(128)
  character(kind=c_char), allocatable, target :: blobs(:)
  integer :: ierr, ok, allok
  allocate(blobs(QNB_IPC_BLOB*numnodes))
  if (qnb_comm_ipc_export(qnb_handle, blob) /= 0) call die('qnb_comm_ipc_export: '//qnb_message())
  call MPI_Allgather(blob, QNB_IPC_BLOB, MPI_BYTE, blobs, QNB_IPC_BLOB, MPI_BYTE, MPI_COMM_WORLD, ierr)
  if (ierr /= 0) call die('qnb_connect_ranks/MPI_Allgather')
  ok = merge(1, 0, qnb_comm_ipc_attach(qnb_handle, nodeid, numnodes, blobs) == 0)
  call MPI_Allreduce(ok, allok, 1, MPI_INTEGER, MPI_MIN, MPI_COMM_WORLD, ierr)
  if (ierr /= 0) call die('qnb_connect_ranks/MPI_Allreduce')
  if (allok == 0) then
     if (nodeid == 0) then
        write(*,'(a)') 'qnb: no peer memory between all GPUs of the job, the sums over the ranks go through NCCL'
        if (qnb_comm_unique_id(id) /= 0) call die('qnb_comm_unique_id: '//qnb_message())
     end if
     call MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr)
     if (ierr /= 0) call die('qnb_connect_ranks/MPI_Bcast')
     if (qnb_comm_init(qnb_handle, nodeid, numnodes, id) /= 0) call die('qnb_comm_init: '//qnb_message())
  end if
  deallocate(blobs)
#else
  call die('USE_QNB: several ranks need the MPI build (-DUSE_MPI)')
#endif
end subroutine qnb_connect_ranks

! --- body of make_pair_lists (nonbondene.f90:749) under USE_QNB
subroutine qnb_glue_make_pair_lists(Rq,Rcq2,RcLRF2,Rcpp2,Rcpw2,Rcww2)
  real(kind=prec) :: Rq,Rcq2,Rcpp2,Rcpw2,Rcww2,RcLRF2
  real(c_double) :: bl(3), ibl(3)
  if (use_PBC) then        ! the box may have changed since the last update (MC_volume, md.f90:1976)
     bl = (/ boxlength%x, boxlength%y, boxlength%z /); ibl = (/ inv_boxl%x, inv_boxl%y, inv_boxl%z /)
     if (qnb_update_box(qnb_handle, bl, ibl) /= 0) call die('qnb_update_box: '//qnb_message())
  end if
  if (qnb_build_lists(qnb_handle, x, real(Rq,c_double), real(Rcq2,c_double), real(RcLRF2,c_double), real(Rcpp2,c_double), &
                      real(Rcpw2,c_double), real(Rcww2,c_double), real(RcLRF,c_double), c_null_ptr) /= 0) &
     call die('make_pair_lists: '//qnb_message())
end subroutine qnb_glue_make_pair_lists

! --- MC_volume (constant_pressure, md.f90:1976-2284) under USE_QNB: three calls added by the patch
! next to "old_nbww = nbww ... old_lrf(:) = lrf(:)" (md.f90:2022-2058): the GPU remembers its last list build
subroutine qnb_glue_save_lists
  if (use_LRF) then
     if (qnb_save_lists(qnb_handle) /= 0) call die('qnb_save_lists: '//qnb_message())
  end if
end subroutine qnb_glue_save_lists

! after boxlength / inv_boxl changed (md.f90:2088-2090).  With LRF on the make_pair_lists that follows (md.f90:2170) sends
! the box again, with LRF off the lists are kept and the trial pot_energy (md.f90:2177) must see the new box
subroutine qnb_glue_update_box
  real(c_double) :: bl(3), ibl(3)
  if (.not. use_PBC) return
  bl = (/ boxlength%x, boxlength%y, boxlength%z /); ibl = (/ inv_boxl%x, inv_boxl%y, inv_boxl%z /)
  if (qnb_update_box(qnb_handle, bl, ibl) /= 0) call die('qnb_update_box: '//qnb_message())
end subroutine qnb_glue_update_box

! volume change rejected (md.f90:2203-2256), called after the host has put boxlength, lists and lrf(:) back
subroutine qnb_glue_restore_lists
  if (use_LRF) then        ! lists, LRF moments and the box of the saved build, bit for bit
     if (qnb_restore_lists(qnb_handle) /= 0) call die('qnb_restore_lists: '//qnb_message())
  else
     call qnb_glue_update_box
  end if
end subroutine qnb_glue_restore_lists

! --- body of pot_energy_nonbonds (potene.f90:320) under USE_QNB; d was zeroed at potene.f90:109 and is added to
subroutine qnb_glue_nonbond(E_loc,EQ_loc,md)
  TYPE(ENERGIES)    :: E_loc
  TYPE(OQ_ENERGIES) :: EQ_loc(:)
  logical           :: md
  real(c_double) :: En(7), lam(max(nstates,1))
  integer :: flags, is
  lam = 0
  lam(1:nstates) = EQ_loc(1:nstates)%lambda
  ! pot_energy has just cleared d (d(:) = zero, potene.f90:109) and the nonbonded terms are its first contribution
  flags = QNB_FLAG_QQ + QNB_FLAG_D_IS_ZERO
  if (md) flags = flags + QNB_FLAG_MD
  if (qnb_nonbond(qnb_handle, x, lam, flags, d, En, qnb_EQ) /= 0) call die('pot_energy_nonbonds: '//qnb_message())
  E_loc%pp%el = E_loc%pp%el + En(1); E_loc%pp%vdw = E_loc%pp%vdw + En(2)
  E_loc%pw%el = E_loc%pw%el + En(3); E_loc%pw%vdw = E_loc%pw%vdw + En(4)
  E_loc%ww%el = E_loc%ww%el + En(5); E_loc%ww%vdw = E_loc%ww%vdw + En(6)
  E_loc%lrf   = E_loc%lrf   + En(7)
  do is = 1, nstates            ! qnb_EQ(6*(is-1)+1:6*is) = qq.el qq.vdw qp.el qp.vdw qw.el qw.vdw
     EQ_loc(is)%qp%el = EQ_loc(is)%qp%el + qnb_EQ(6*(is-1)+3); EQ_loc(is)%qp%vdw = EQ_loc(is)%qp%vdw + qnb_EQ(6*(is-1)+4)
     EQ_loc(is)%qw%el = EQ_loc(is)%qw%el + qnb_EQ(6*(is-1)+5); EQ_loc(is)%qw%vdw = EQ_loc(is)%qw%vdw + qnb_EQ(6*(is-1)+6)
  end do
end subroutine qnb_glue_nonbond

! --- body of nonbond_qq (potene.f90:176) under USE_QNB: the static-list terms were evaluated in the same GPU call, the
!     gradient is already in d, only EQ%qq is left to add; nonbond_qqp (L177) has nothing left to do (its energies are
!     part of qp above, as in the reference)
subroutine qnb_glue_add_qq(EQ_loc)
  TYPE(NB_ENERGIES) :: EQ_loc(:)
  integer :: is
  do is = 1, nstates
     EQ_loc(is)%el = EQ_loc(is)%el + qnb_EQ(6*(is-1)+1); EQ_loc(is)%vdw = EQ_loc(is)%vdw + qnb_EQ(6*(is-1)+2)
  end do
end subroutine qnb_glue_add_qq

subroutine qnb_shutdown
  integer(c_int) :: rc
  if (c_associated(qnb_handle)) rc = qnb_finalize(qnb_handle)
  qnb_handle = c_null_ptr
end subroutine qnb_shutdown

end module QNB_GLUE
