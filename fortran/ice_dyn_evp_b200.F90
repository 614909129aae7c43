!=======================================================================
! ice_dyn_evp_b200 -- Fortran side of the B200 EVP path (ISO_C_BINDING shim over include/evp_b200.h)
!
! NOT COMPILED OR TESTED IN THIS REPOSITORY'S IMAGE: there is no Fortran compiler here (gcc lacks
! f951, no MPI).  tests/ exercise the same C ABI from a caller that reproduces this module's memory
! layout byte for byte (Fortran a(nx_block,ny_block,max_blocks) == C-ordered (max_blocks,ny_block,nx_block)).
!
! It plugs into CICE at the seam the reference already has for its own 1-D solver:
!   dyn_evp_b200_init      next to  dyn_evp1d_init      (ice_dyn_evp.F90:153-155)
!   dyn_evp_b200_run       next to  dyn_evp1d_run       (ice_dyn_evp.F90:846-858), same 31 arrays
!   dyn_evp_b200_finalize  next to  dyn_evp1d_finalize
! and is selected with  evp_algorithm = 'gpu_b200'  in dynamics_nml (see INTEGRATION.md for the patch).
!=======================================================================
module ice_dyn_evp_b200

  use, intrinsic :: iso_c_binding
  use ice_kinds_mod
  use ice_blocks,      only: block, get_block, nx_block, ny_block, nghost
  use ice_domain,      only: nblocks, blocks_ice, ew_boundary_type, ns_boundary_type
  use ice_domain_size, only: max_blocks, nx_global, ny_global
  use ice_communicate, only: my_task, master_task, MPI_COMM_ICE, get_num_procs
  use ice_exit,        only: abort_ice

  implicit none
  private
  public :: dyn_evp_b200_init, dyn_evp_b200_run, dyn_evp_b200_finalize
  public :: dyn_evp_b200_init_cgrid, dyn_evp_b200_run_cgrid   ! grid_ice = 'C' (ice_dyn_evp.F90:936-1101)
  public :: dyn_evp_b200_deformations, dyn_evp_b200_finish    ! the two steps right after the loop, from the device-resident velocities
  public :: dyn_evp_b200_prep_init, dyn_evp_b200_step         ! step preparation on the device, dynamics state resident (SURVEY 8f ranks 1, 3)

  integer(c_int32_t), parameter :: EVP_B200_ABI_VERSION = 3
  integer(c_int32_t), parameter :: BNDY_OPEN = 0, BNDY_CLOSED = 1, BNDY_CYCLIC = 2, BNDY_TRIPOLE = 3

  ! evp_b200_grid_t
  type, bind(C) :: evp_b200_grid_t
     integer(c_int32_t) :: abi_version, nx_block, ny_block, nblocks, max_blocks, nghost
     integer(c_int32_t) :: nx_global, ny_global, ew_boundary_type, ns_boundary_type
     type(c_ptr) :: ilo, ihi, jlo, jhi, i_glob, j_glob
     type(c_ptr) :: dxT, dyT, dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea, uarear
  end type evp_b200_grid_t

  ! evp_b200_params_t
  type, bind(C) :: evp_b200_params_t
     integer(c_int32_t) :: ndte, mode, kernel, visc_method
     real(c_double) :: arlx1i, denom1, revp, brlx, e_factor, epp2i, capping, Ktens, u0, cosw, sinw, rhow, deltaminEVP
  end type evp_b200_params_t

  ! evp_b200_fields_t: the argument list of dyn_evp1d_run, in order
  type, bind(C) :: evp_b200_fields_t
     type(c_ptr) :: stressp_1, stressp_2, stressp_3, stressp_4
     type(c_ptr) :: stressm_1, stressm_2, stressm_3, stressm_4
     type(c_ptr) :: stress12_1, stress12_2, stress12_3, stress12_4
     type(c_ptr) :: strength, cdn_ocnU, aiU, uocnU, vocnU, waterxU, wateryU, forcexU, forceyU, umassdti, fmU
     type(c_ptr) :: strintxU, strintyU, TbU, taubxU, taubyU, uvel, vvel
     type(c_ptr) :: iceTmask, iceUmask
  end type evp_b200_fields_t

  ! evp_b200_cgrid_t: extra static geometry of grid_ice = 'C'
  type, bind(C) :: evp_b200_cgrid_t
     type(c_ptr) :: dxN, dyE, dxE, dyN, dxU, dyU
     type(c_ptr) :: tarea, uarea, earea, narea, earear, narear
     type(c_ptr) :: ratiodxN, ratiodxNr, ratiodyE, ratiodyEr
     type(c_ptr) :: hm, uvm, epm, npm
  end type evp_b200_cgrid_t

  ! evp_b200_cfields_t: what the C-grid subcycle loop reads and writes, in the order of include/evp_b200.h
  type, bind(C) :: evp_b200_cfields_t
     type(c_ptr) :: uvelE, vvelE, uvelN, vvelN, uvel, vvel
     type(c_ptr) :: stresspT, stressmT, stress12T, stress12U
     type(c_ptr) :: zetax2T, etax2T, etax2U, strengthU
     type(c_ptr) :: divergU, tensionU, shearU, deltaU
     type(c_ptr) :: strintxE, strintyN, taubxE, taubyN
     type(c_ptr) :: strength
     type(c_ptr) :: cdn_ocnE, cdn_ocnN, aiE, aiN, uocnE, vocnE, uocnN, vocnN
     type(c_ptr) :: waterxE, wateryN, forcexE, forceyN, emassdti, nmassdti, fmE, fmN
     type(c_ptr) :: TbE, TbN, rheofactE, rheofactN
     type(c_ptr) :: iceTmask, iceUmask, iceEmask, iceNmask
  end type evp_b200_cfields_t

  ! evp_b200_deform_t
  type, bind(C) :: evp_b200_deform_t
     type(c_ptr) :: dxU, dyU, tarear
     type(c_ptr) :: divu, shear, vort, rdg_conv, rdg_shear
     real(c_double) :: e_factor
  end type evp_b200_deform_t

  ! evp_b200_finish_t
  type, bind(C) :: evp_b200_finish_t
     type(c_ptr) :: strocnxU, strocnyU
     real(c_double) :: rhow, cosw, sinw
  end type evp_b200_finish_t

  ! evp_b200_prep_static_t, evp_b200_prep_t (include/evp_b200.h)
  type, bind(C) :: evp_b200_prep_static_t
     type(c_ptr) :: hm, tarea, uarea, fcor, umask
  end type evp_b200_prep_static_t
  type, bind(C) :: evp_b200_prep_t
     type(c_ptr) :: tmass, aice_init, cdn_ocn, uocn, vocn, ss_tltx, ss_tlty, strairxT, strairyT, strength
     type(c_ptr) :: iceTmask, TbU
     real(c_double) :: dt, dyn_area_min, dyn_mass_min, gravit
     integer(c_int32_t) :: ssh_stress
  end type evp_b200_prep_t

  interface
     integer(c_int) function evp_b200_prep_init(st) bind(C, name='evp_b200_prep_init')
       import :: c_int, evp_b200_prep_static_t
       type(evp_b200_prep_static_t), intent(in) :: st
     end function evp_b200_prep_init
     integer(c_int) function evp_b200_step_resident(params, prep, fields, flags) bind(C, name='evp_b200_step_resident')
       import :: c_int, c_int32_t, evp_b200_params_t, evp_b200_prep_t, evp_b200_fields_t
       type(evp_b200_params_t), intent(in) :: params
       type(evp_b200_prep_t), intent(in) :: prep
       type(evp_b200_fields_t), intent(inout) :: fields
       integer(c_int32_t), value :: flags
     end function evp_b200_step_resident
     integer(c_int) function evp_b200_pin_host(ptr, bytes) bind(C, name='evp_b200_pin_host')
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ptr
       integer(c_size_t), value :: bytes
     end function evp_b200_pin_host
     integer(c_int) function evp_b200_set_metric(HTN, HTE, deltaminEVP, mismatches) bind(C, name='evp_b200_set_metric')
       import :: c_int, c_int32_t, c_double, c_ptr
       type(c_ptr), value :: HTN, HTE
       real(c_double), value :: deltaminEVP
       integer(c_int32_t), intent(out) :: mismatches
     end function evp_b200_set_metric
     integer(c_int) function evp_b200_deformations(d) bind(C, name='evp_b200_deformations')
       import :: c_int, evp_b200_deform_t
       type(evp_b200_deform_t), intent(inout) :: d
     end function evp_b200_deformations
     integer(c_int) function evp_b200_dyn_finish(f) bind(C, name='evp_b200_dyn_finish')
       import :: c_int, evp_b200_finish_t
       type(evp_b200_finish_t), intent(inout) :: f
     end function evp_b200_dyn_finish
     integer(c_int) function evp_b200_init_cgrid(cgrid) bind(C, name='evp_b200_init_cgrid')
       import :: c_int, evp_b200_cgrid_t
       type(evp_b200_cgrid_t), intent(in) :: cgrid
     end function
     integer(c_int) function evp_b200_run_cgrid(params, fields) bind(C, name='evp_b200_run_cgrid')
       import :: c_int, evp_b200_params_t, evp_b200_cfields_t
       type(evp_b200_params_t), intent(in) :: params
       type(evp_b200_cfields_t), intent(inout) :: fields
     end function
     integer(c_int) function evp_b200_get_unique_id(id) bind(C, name='evp_b200_get_unique_id')
       import :: c_int, c_char
       character(kind=c_char) :: id(128)
     end function
     integer(c_int) function evp_b200_comm_init(rank, nranks, id) bind(C, name='evp_b200_comm_init')
       import :: c_int, c_int32_t, c_char
       integer(c_int32_t), value :: rank, nranks
       character(kind=c_char) :: id(128)
     end function
     integer(c_int) function evp_b200_set_device(dev) bind(C, name='evp_b200_set_device')
       import :: c_int, c_int32_t
       integer(c_int32_t), value :: dev
     end function
     integer(c_int) function evp_b200_allow_partial_domain(yes) bind(C, name='evp_b200_allow_partial_domain')
       import :: c_int, c_int32_t
       integer(c_int32_t), value :: yes
     end function
     integer(c_int) function evp_b200_init(grid) bind(C, name='evp_b200_init')
       import :: c_int, evp_b200_grid_t
       type(evp_b200_grid_t), intent(in) :: grid
     end function
     integer(c_int) function evp_b200_run_bgrid(params, fields) bind(C, name='evp_b200_run_bgrid')
       import :: c_int, evp_b200_params_t, evp_b200_fields_t
       type(evp_b200_params_t), intent(in) :: params
       type(evp_b200_fields_t), intent(inout) :: fields
     end function
     integer(c_int) function evp_b200_run_bgrid_resident(params, fields, flags) bind(C, name='evp_b200_run_bgrid_resident')
       import :: c_int, c_int32_t, evp_b200_params_t, evp_b200_fields_t
       type(evp_b200_params_t), intent(in) :: params
       type(evp_b200_fields_t), intent(inout) :: fields
       integer(c_int32_t), value :: flags
     end function
     integer(c_int) function evp_b200_download_stress(fields) bind(C, name='evp_b200_download_stress')
       import :: c_int, evp_b200_fields_t
       type(evp_b200_fields_t), intent(inout) :: fields
     end function
     ! tripole grids, stresses resident: the twelve ice_HaloUpdate_stress calls of ice_dyn_evp.F90:1321-1388 on the device
     ! (evp_b200_run_bgrid_resident calls it itself; for users of the upload / subcycle / download split)
     integer(c_int) function evp_b200_stress_symmetrise() bind(C, name='evp_b200_stress_symmetrise')
       import :: c_int
     end function
     integer(c_int) function evp_b200_finalize() bind(C, name='evp_b200_finalize')
       import :: c_int
     end function
     type(c_ptr) function evp_b200_last_error() bind(C, name='evp_b200_last_error')
       import :: c_ptr
     end function
  end interface

  ! C-interoperable copies of the block table and of the logical masks
  integer(c_int32_t), allocatable, target, save :: b_ilo(:), b_ihi(:), b_jlo(:), b_jhi(:)
  integer(c_int32_t), allocatable, target, save :: b_iglob(:,:), b_jglob(:,:)
  integer(c_int32_t), allocatable, target, save :: imaskT(:,:,:), imaskU(:,:,:)
  integer(c_int32_t), allocatable, target, save :: imaskE(:,:,:), imaskN(:,:,:)   ! C grid
  integer(c_int32_t), allocatable, target, save :: iumask(:,:,:)                  ! step preparation: umask as 0/1
  logical, save :: pinned = .false.   ! the B-grid field arrays have been page-locked (evp_b200_pin_host)

contains

  !---------------------------------------------------------------------
  subroutine check(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: text
    integer :: n
    if (rc == 0) return
    call c_f_pointer(evp_b200_last_error(), msg, [512])
    text = ' '
    do n = 1, 512
       if (msg(n) == c_null_char) exit
       text(n:n) = msg(n)
    enddo
    ! error convention of the reference: comm/mpi/ice_exit.F90
    call abort_ice('(ice_dyn_evp_b200) ERROR in '//trim(where)//': '//trim(text), file=__FILE__, line=__LINE__)
  end subroutine check

  integer(c_int32_t) function bndy_code(name)
    character(len=*), intent(in) :: name
    select case (trim(name))
    case ('cyclic');  bndy_code = BNDY_CYCLIC
    case ('closed');  bndy_code = BNDY_CLOSED
    case ('tripole'); bndy_code = BNDY_TRIPOLE
    case default;     bndy_code = BNDY_OPEN
    end select
    if (trim(name) == 'tripoleT') call abort_ice('(ice_dyn_evp_b200) ERROR: tripoleT not supported', &
         file=__FILE__, line=__LINE__)
  end function bndy_code

  !---------------------------------------------------------------------
  ! once, after init_dyn_shared and the grid are set up (ice_dyn_evp.F90:153-155)
  subroutine dyn_evp_b200_init
    use mpi
    use ice_grid,       only: dxT, dyT, uarear, HTN, HTE
    use ice_dyn_shared, only: dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea, deltaminEVP
    type(evp_b200_grid_t) :: g
    type(block) :: this_block
    character(kind=c_char) :: id(128)
    integer(c_int32_t) :: nbad
    integer :: iblk, ierr, nprocs, ndev_local
    integer :: local_comm, local_rank

    ! one rank <-> one GPU of the node
    call MPI_Comm_split_type(MPI_COMM_ICE, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, local_comm, ierr)
    call MPI_Comm_rank(local_comm, local_rank, ierr)
    call check(evp_b200_set_device(int(local_rank, c_int32_t)), 'evp_b200_set_device')

    ! NCCL bootstrap over the host's own transport (replaces the MPI halo of ice_boundary for the dyn fields)
    nprocs = get_num_procs()
    if (nprocs > 1) then
       if (my_task == master_task) call check(evp_b200_get_unique_id(id), 'evp_b200_get_unique_id')
       call MPI_Bcast(id, 128, MPI_CHARACTER, master_task, MPI_COMM_ICE, ierr)
       call check(evp_b200_comm_init(int(my_task, c_int32_t), int(nprocs, c_int32_t), id), 'evp_b200_comm_init')
    else
       ! one task holds every distributed block: if they span less than the domain, land blocks were eliminated (ice_domain.F90)
       call check(evp_b200_allow_partial_domain(1_c_int32_t), 'evp_b200_allow_partial_domain')
    endif

    allocate(b_ilo(nblocks), b_ihi(nblocks), b_jlo(nblocks), b_jhi(nblocks))
    allocate(b_iglob(nx_block, nblocks), b_jglob(ny_block, nblocks))
    allocate(imaskT(nx_block, ny_block, max_blocks), imaskU(nx_block, ny_block, max_blocks))
    do iblk = 1, nblocks
       this_block = get_block(blocks_ice(iblk), iblk)
       b_ilo(iblk) = this_block%ilo;  b_ihi(iblk) = this_block%ihi
       b_jlo(iblk) = this_block%jlo;  b_jhi(iblk) = this_block%jhi
       b_iglob(:, iblk) = this_block%i_glob(:)
       b_jglob(:, iblk) = this_block%j_glob(:)
    enddo

    g%abi_version = EVP_B200_ABI_VERSION
    g%nx_block = nx_block;  g%ny_block = ny_block;  g%nblocks = nblocks;  g%max_blocks = max_blocks
    g%nghost = nghost;      g%nx_global = nx_global; g%ny_global = ny_global
    g%ew_boundary_type = bndy_code(ew_boundary_type)
    g%ns_boundary_type = bndy_code(ns_boundary_type)
    g%ilo = c_loc(b_ilo);  g%ihi = c_loc(b_ihi);  g%jlo = c_loc(b_jlo);  g%jhi = c_loc(b_jhi)
    g%i_glob = c_loc(b_iglob);  g%j_glob = c_loc(b_jglob)
    g%dxT = loc3(dxT);    g%dyT = loc3(dyT);    g%dxhy = loc3(dxhy);  g%dyhx = loc3(dyhx)
    g%cxp = loc3(cxp);    g%cyp = loc3(cyp);    g%cxm = loc3(cxm);    g%cym = loc3(cym)
    g%DminTarea = loc3(DminTarea);  g%uarear = loc3(uarear)
    call check(evp_b200_init(g), 'evp_b200_init')
    ! optional: the metric arrays behind dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea; the library checks on the device that they
    ! reproduce those arrays bit for bit before any kernel may derive them (include/evp_b200.h); a non-zero count is not an error
    call check(evp_b200_set_metric(loc3(HTN), loc3(HTE), deltaminEVP, nbad), 'evp_b200_set_metric')
  end subroutine dyn_evp_b200_init

  !---------------------------------------------------------------------
  ! one dynamics step: replaces the `do ksub = 1,ndte` loop of ice_dyn_evp.F90:859-913.
  ! Same argument list and order as dyn_evp1d_run (ice_dyn_evp1d.F90:119-153).
  subroutine dyn_evp_b200_run(stressp_1 , stressp_2 , stressp_3 , stressp_4 , &
                              stressm_1 , stressm_2 , stressm_3 , stressm_4 , &
                              stress12_1, stress12_2, stress12_3, stress12_4, &
                              strength  ,                                     &
                              cdn_ocnU  , aiU       , uocnU     , vocnU     , &
                              waterxU   , wateryU   , forcexU   , forceyU   , &
                              umassdti  , fmU       , strintxU  , strintyU  , &
                              TbU       , taubxU    , taubyU    , uvel      , &
                              vvel      , iceTmask  , iceUmask  , stress_on_host)
    use ice_dyn_shared, only: ndte, arlx1i, denom1, revp, brlx, e_factor, epp2i, capping, Ktens, u0, cosw, sinw, deltaminEVP
    use icepack_intfc,  only: icepack_query_parameters
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: &
         stressp_1 , stressp_2 , stressp_3 , stressp_4 , stressm_1 , stressm_2 , stressm_3 , stressm_4 , &
         stress12_1, stress12_2, stress12_3, stress12_4, strintxU  , strintyU  , uvel      , vvel      , &
         taubxU    , taubyU
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous :: &
         strength, cdn_ocnU, aiU, uocnU, vocnU, waterxU, wateryU, forcexU, forceyU, umassdti, fmU, TbU
    logical(kind=log_kind), dimension(:,:,:), intent(in) :: iceTmask, iceUmask
    ! optional: .false. keeps the 12 stress arrays on the device between steps (EVP_B200_KEEP_STRESS, include/evp_b200.h);
    ! the caller passes .true. on steps that write a restart or history file (ice_restart_driver.F90:150-231).
    ! Absent = the host arrays are read and written every step, exactly like dyn_evp1d_run.
    logical(kind=log_kind), intent(in), optional :: stress_on_host
    integer(c_int32_t) :: flags
    type(evp_b200_params_t) :: p
    type(evp_b200_fields_t) :: f
    real(kind=dbl_kind) :: rhow

    call icepack_query_parameters(rhow_out=rhow)
    p%ndte = ndte;  p%mode = 0;  p%kernel = 0;  p%visc_method = 0   ! exact arithmetic, library picks the kernel
    p%arlx1i = arlx1i;  p%denom1 = denom1;  p%revp = revp;  p%brlx = brlx
    p%e_factor = e_factor;  p%epp2i = epp2i;  p%capping = capping;  p%Ktens = Ktens
    p%u0 = u0;  p%cosw = cosw;  p%sinw = sinw;  p%rhow = rhow;  p%deltaminEVP = deltaminEVP

    ! Fortran logical is not C-interoperable: 0/1 integers cross the boundary
    imaskT = merge(1_c_int32_t, 0_c_int32_t, iceTmask)
    imaskU = merge(1_c_int32_t, 0_c_int32_t, iceUmask)

    f%stressp_1 = c_loc(stressp_1);   f%stressp_2 = c_loc(stressp_2);   f%stressp_3 = c_loc(stressp_3);   f%stressp_4 = c_loc(stressp_4)
    f%stressm_1 = c_loc(stressm_1);   f%stressm_2 = c_loc(stressm_2);   f%stressm_3 = c_loc(stressm_3);   f%stressm_4 = c_loc(stressm_4)
    f%stress12_1 = c_loc(stress12_1); f%stress12_2 = c_loc(stress12_2); f%stress12_3 = c_loc(stress12_3); f%stress12_4 = c_loc(stress12_4)
    f%strength = c_loc(strength);  f%cdn_ocnU = c_loc(cdn_ocnU);  f%aiU = c_loc(aiU)
    f%uocnU = c_loc(uocnU);        f%vocnU = c_loc(vocnU)
    f%waterxU = c_loc(waterxU);    f%wateryU = c_loc(wateryU);    f%forcexU = c_loc(forcexU);  f%forceyU = c_loc(forceyU)
    f%umassdti = c_loc(umassdti);  f%fmU = c_loc(fmU)
    f%strintxU = c_loc(strintxU);  f%strintyU = c_loc(strintyU);  f%TbU = c_loc(TbU)
    f%taubxU = c_loc(taubxU);      f%taubyU = c_loc(taubyU);      f%uvel = c_loc(uvel);        f%vvel = c_loc(vvel)
    f%iceTmask = c_loc(imaskT);    f%iceUmask = c_loc(imaskU)

    ! the arrays are module variables of ice_dyn_evp / ice_flux / ice_state and never move: page-lock them on the first call, so
    ! that every later copy runs at the PCIe rate instead of through the driver's staging of pageable memory
    if (.not. pinned) then
       call pin(stressp_1);  call pin(stressp_2);  call pin(stressp_3);  call pin(stressp_4)
       call pin(stressm_1);  call pin(stressm_2);  call pin(stressm_3);  call pin(stressm_4)
       call pin(stress12_1); call pin(stress12_2); call pin(stress12_3); call pin(stress12_4)
       call pin(strength);   call pin(cdn_ocnU);   call pin(aiU);        call pin(uocnU);     call pin(vocnU)
       call pin(waterxU);    call pin(wateryU);    call pin(forcexU);    call pin(forceyU);   call pin(umassdti);  call pin(fmU)
       call pin(strintxU);   call pin(strintyU);   call pin(TbU);        call pin(taubxU);    call pin(taubyU)
       call pin(uvel);       call pin(vvel)
       call check(evp_b200_pin_host(c_loc(imaskT), int(size(imaskT), c_size_t) * 4_c_size_t), 'evp_b200_pin_host')
       call check(evp_b200_pin_host(c_loc(imaskU), int(size(imaskU), c_size_t) * 4_c_size_t), 'evp_b200_pin_host')
       pinned = .true.
    endif

    flags = 0
    if (present(stress_on_host)) then
       flags = 1                               ! EVP_B200_KEEP_STRESS
       if (stress_on_host) flags = 3           ! ... | EVP_B200_FETCH_STRESS
    endif
    if (flags /= 0) then
       call check(evp_b200_run_bgrid_resident(p, f, flags), 'evp_b200_run_bgrid_resident')
    else
       call check(evp_b200_run_bgrid(p, f), 'evp_b200_run_bgrid')
    endif
  end subroutine dyn_evp_b200_run

  subroutine pin(a)
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous :: a
    call check(evp_b200_pin_host(c_loc(a), int(size(a), c_size_t) * 8_c_size_t), 'evp_b200_pin_host')
  end subroutine pin

  ! C address of a use-associated module array.  The reference declares its grid and dynamics arrays without TARGET
  ! (ice_grid.F90:78-140, ice_dyn_shared.F90:100-130), and C_LOC requires POINTER or TARGET (gfortran: "shall have either the
  ! POINTER or the TARGET attribute"), so the address is taken of a TARGET, CONTIGUOUS dummy instead: the arrays are whole
  ! allocatables, hence contiguous, hence passed by reference without copy-in, and the address stays valid for the run.
  function loc3(a) result(p)
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous :: a
    type(c_ptr) :: p
    p = c_loc(a)
  end function loc3

  !---------------------------------------------------------------------
  ! `deformations` (ice_dyn_shared.F90:1756-1860; the per-block call at ice_dyn_evp.F90:920-934) for all blocks at once, from
  ! the velocities dyn_evp_b200_run left on the device.  The five arrays keep their values off the ice T cells.
  subroutine dyn_evp_b200_deformations(divu, shear, vort, rdg_conv, rdg_shear)
    use ice_grid,       only: dxU, dyU, tarear
    use ice_dyn_shared, only: e_factor
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: divu, shear, vort, rdg_conv, rdg_shear
    type(evp_b200_deform_t) :: d
    d%dxU = loc3(dxU);  d%dyU = loc3(dyU);  d%tarear = loc3(tarear)
    d%divu = c_loc(divu);  d%shear = c_loc(shear);  d%vort = c_loc(vort);  d%rdg_conv = c_loc(rdg_conv);  d%rdg_shear = c_loc(rdg_shear)
    d%e_factor = e_factor
    call check(evp_b200_deformations(d), 'evp_b200_deformations')
  end subroutine dyn_evp_b200_deformations

  !---------------------------------------------------------------------
  ! `dyn_finish` (ice_dyn_shared.F90:1291-1365; the per-block call at ice_dyn_evp.F90:1392-1405) for all blocks at once: the
  ! velocities, cdn_ocnU, uocnU, vocnU, aiU, fmU and iceUmask of the last dyn_evp_b200_run are still on the device.
  subroutine dyn_evp_b200_finish(strocnxU, strocnyU)
    use ice_dyn_shared, only: cosw, sinw
    use icepack_intfc,  only: icepack_query_parameters
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: strocnxU, strocnyU
    type(evp_b200_finish_t) :: f
    real(kind=dbl_kind) :: rhow
    call icepack_query_parameters(rhow_out=rhow)
    f%strocnxU = c_loc(strocnxU);  f%strocnyU = c_loc(strocnyU)
    f%rhow = rhow;  f%cosw = cosw;  f%sinw = sinw
    call check(evp_b200_dyn_finish(f), 'evp_b200_dyn_finish')
  end subroutine dyn_evp_b200_finish

  !---------------------------------------------------------------------
  ! Step preparation on the device (SURVEY 8f ranks 1 and 3; include/evp_b200.h: evp_b200_prep_init / evp_b200_step_resident).
  ! Once, after dyn_evp_b200_init: the static inputs of grid_average_X2Y 'S'/'F' (ice_grid.F90:4159-4211, 4620-4660) and of dyn_prep2.
  subroutine dyn_evp_b200_prep_init
    use ice_grid,       only: hm, tarea, uarea, umask
    use ice_dyn_shared, only: fcor_blk
    type(evp_b200_prep_static_t) :: st
    allocate(iumask(nx_block, ny_block, max_blocks))
    iumask = merge(1_c_int32_t, 0_c_int32_t, umask)
    st%hm = loc3(hm);  st%tarea = loc3(tarea);  st%uarea = loc3(uarea);  st%fcor = loc3(fcor_blk);  st%umask = c_loc(iumask)
    call check(evp_b200_prep_init(st), 'evp_b200_prep_init')
  end subroutine dyn_evp_b200_prep_init

  ! One dynamics step: replaces ice_dyn_evp.F90:428-531 (T -> U averages, dyn_prep2), :735-739 (velocity halo update) and the
  ! subcycle loop :859-913.  The T-point arrays have had their halo updates (:419-426, 471-474), strength its own (:731-733).
  ! Velocities, stresses and iceUmask live on the device; uvel, vvel are returned every step.
  subroutine dyn_evp_b200_step(tmass, aice_init, cdn_ocn, uocn, vocn, ss_tltx, ss_tlty, strairxT, strairyT, strength, iceTmask, &
                               uvel, vvel, first_step, want_diag, want_state,                                                    &
                               stressp_1 , stressp_2 , stressp_3 , stressp_4 , stressm_1 , stressm_2 , stressm_3 , stressm_4 ,   &
                               stress12_1, stress12_2, stress12_3, stress12_4, strintxU, strintyU, taubxU, taubyU, iceUmask, TbU)
    use ice_dyn_shared, only: ndte, arlx1i, denom1, revp, brlx, e_factor, epp2i, capping, Ktens, u0, cosw, sinw, deltaminEVP, &
                              dyn_area_min, dyn_mass_min, ssh_stress
    use ice_calendar,   only: dt_dyn
    use icepack_intfc,  only: icepack_query_parameters
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous :: &
         tmass, aice_init, cdn_ocn, uocn, vocn, ss_tltx, ss_tlty, strairxT, strairyT, strength
    logical(kind=log_kind), dimension(:,:,:), intent(in) :: iceTmask
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: uvel, vvel
    logical(kind=log_kind), intent(in) :: first_step, want_diag, want_state
    ! carried state and diagnostics: read when first_step, written when want_state / want_diag
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: &
         stressp_1 , stressp_2 , stressp_3 , stressp_4 , stressm_1 , stressm_2 , stressm_3 , stressm_4 , &
         stress12_1, stress12_2, stress12_3, stress12_4, strintxU, strintyU, taubxU, taubyU
    logical(kind=log_kind), dimension(:,:,:), intent(inout) :: iceUmask
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous, optional :: TbU   ! seabed_stress = .true. only
    type(evp_b200_params_t) :: p
    type(evp_b200_prep_t)   :: pr
    type(evp_b200_fields_t) :: f
    integer(c_int32_t) :: flags
    real(kind=dbl_kind) :: rhow, gravit

    call icepack_query_parameters(rhow_out=rhow, gravit_out=gravit)
    p%ndte = ndte;  p%mode = 0;  p%kernel = 0;  p%visc_method = 0
    p%arlx1i = arlx1i;  p%denom1 = denom1;  p%revp = revp;  p%brlx = brlx
    p%e_factor = e_factor;  p%epp2i = epp2i;  p%capping = capping;  p%Ktens = Ktens
    p%u0 = u0;  p%cosw = cosw;  p%sinw = sinw;  p%rhow = rhow;  p%deltaminEVP = deltaminEVP

    imaskT = merge(1_c_int32_t, 0_c_int32_t, iceTmask)
    pr%tmass = c_loc(tmass);  pr%aice_init = c_loc(aice_init);  pr%cdn_ocn = c_loc(cdn_ocn)
    pr%uocn = c_loc(uocn);    pr%vocn = c_loc(vocn);            pr%ss_tltx = c_loc(ss_tltx);  pr%ss_tlty = c_loc(ss_tlty)
    pr%strairxT = c_loc(strairxT);  pr%strairyT = c_loc(strairyT);  pr%strength = c_loc(strength)
    pr%iceTmask = c_loc(imaskT)
    pr%TbU = c_null_ptr
    if (present(TbU)) pr%TbU = c_loc(TbU)
    pr%dt = dt_dyn;  pr%dyn_area_min = dyn_area_min;  pr%dyn_mass_min = dyn_mass_min;  pr%gravit = gravit
    pr%ssh_stress = 0
    if (trim(ssh_stress) == 'coupled') pr%ssh_stress = 1

    f%stressp_1 = c_loc(stressp_1);   f%stressp_2 = c_loc(stressp_2);   f%stressp_3 = c_loc(stressp_3);   f%stressp_4 = c_loc(stressp_4)
    f%stressm_1 = c_loc(stressm_1);   f%stressm_2 = c_loc(stressm_2);   f%stressm_3 = c_loc(stressm_3);   f%stressm_4 = c_loc(stressm_4)
    f%stress12_1 = c_loc(stress12_1); f%stress12_2 = c_loc(stress12_2); f%stress12_3 = c_loc(stress12_3); f%stress12_4 = c_loc(stress12_4)
    f%strintxU = c_loc(strintxU);  f%strintyU = c_loc(strintyU);  f%taubxU = c_loc(taubxU);  f%taubyU = c_loc(taubyU)
    f%uvel = c_loc(uvel);  f%vvel = c_loc(vvel)
    imaskU = merge(1_c_int32_t, 0_c_int32_t, iceUmask)
    f%iceTmask = c_loc(imaskT);  f%iceUmask = c_loc(imaskU)

    flags = 0
    if (first_step) flags = flags + 1          ! EVP_B200_STEP_INIT_STATE
    if (want_diag)  flags = flags + 2          ! EVP_B200_STEP_FETCH_DIAG
    if (want_state) flags = flags + 4          ! EVP_B200_STEP_FETCH_STATE
    call check(evp_b200_step_resident(p, pr, f, flags), 'evp_b200_step_resident')
    if (want_state) iceUmask = (imaskU /= 0)
  end subroutine dyn_evp_b200_step

  !---------------------------------------------------------------------
  ! grid_ice = 'C': once, after dyn_evp_b200_init (the ratio arrays exist after init_evp, ice_dyn_evp.F90:218-240)
  subroutine dyn_evp_b200_init_cgrid
    use ice_grid,    only: dxN, dyE, dxE, dyN, dxU, dyU, tarea, uarea, earea, narea, earear, narear, hm, uvm, epm, npm
    use ice_dyn_evp, only: ratiodxN, ratiodxNr, ratiodyE, ratiodyEr
    type(evp_b200_cgrid_t) :: cg
    cg%dxN = loc3(dxN);  cg%dyE = loc3(dyE);  cg%dxE = loc3(dxE);  cg%dyN = loc3(dyN);  cg%dxU = loc3(dxU);  cg%dyU = loc3(dyU)
    cg%tarea = loc3(tarea);  cg%uarea = loc3(uarea);  cg%earea = loc3(earea);  cg%narea = loc3(narea)
    cg%earear = loc3(earear);  cg%narear = loc3(narear)
    cg%ratiodxN = loc3(ratiodxN);  cg%ratiodxNr = loc3(ratiodxNr);  cg%ratiodyE = loc3(ratiodyE);  cg%ratiodyEr = loc3(ratiodyEr)
    cg%hm = loc3(hm);  cg%uvm = loc3(uvm);  cg%epm = loc3(epm);  cg%npm = loc3(npm)
    allocate(imaskE(nx_block, ny_block, max_blocks), imaskN(nx_block, ny_block, max_blocks))
    call check(evp_b200_init_cgrid(cg), 'evp_b200_init_cgrid')
  end subroutine dyn_evp_b200_init_cgrid

  !---------------------------------------------------------------------
  ! grid_ice = 'C': one dynamics step, replaces the `do ksub = 1,ndte` loop of ice_dyn_evp.F90:938-1097 (one more
  ! elseif in front of it).  The arguments are the module / local arrays that loop reads and writes.
  subroutine dyn_evp_b200_run_cgrid(uvelE, vvelE, uvelN, vvelN, uvel, vvel,                          &
                                    stresspT, stressmT, stress12T, stress12U,                        &
                                    zetax2T, etax2T, etax2U, strengthU,                              &
                                    divergU, tensionU, shearU, deltaU,                               &
                                    strintxE, strintyN, taubxE, taubyN,                              &
                                    strength,                                                        &
                                    cdn_ocnE, cdn_ocnN, aiE, aiN, uocnE, vocnE, uocnN, vocnN,        &
                                    waterxE, wateryN, forcexE, forceyN, emassdti, nmassdti, fmE, fmN,&
                                    TbE, TbN, rheofactE, rheofactN,                                  &
                                    iceTmask, iceUmask, iceEmask, iceNmask)
    use ice_dyn_shared, only: ndte, arlx1i, denom1, revp, brlx, e_factor, epp2i, capping, Ktens, u0, cosw, sinw, deltaminEVP, &
                              visc_method
    use icepack_intfc,  only: icepack_query_parameters
    real(kind=dbl_kind), dimension(:,:,:), intent(inout), target, contiguous :: &
         uvelE, vvelE, uvelN, vvelN, uvel, vvel, stresspT, stressmT, stress12T, stress12U, &
         zetax2T, etax2T, etax2U, strengthU, divergU, tensionU, shearU, deltaU, strintxE, strintyN, taubxE, taubyN
    real(kind=dbl_kind), dimension(:,:,:), intent(in), target, contiguous :: &
         strength, cdn_ocnE, cdn_ocnN, aiE, aiN, uocnE, vocnE, uocnN, vocnN, waterxE, wateryN, forcexE, forceyN, &
         emassdti, nmassdti, fmE, fmN, TbE, TbN, rheofactE, rheofactN
    logical(kind=log_kind), dimension(:,:,:), intent(in) :: iceTmask, iceUmask, iceEmask, iceNmask
    type(evp_b200_params_t) :: p
    type(evp_b200_cfields_t) :: f
    real(kind=dbl_kind) :: rhow

    call icepack_query_parameters(rhow_out=rhow)
    p%ndte = ndte;  p%mode = 0;  p%kernel = 0
    p%visc_method = merge(1_c_int32_t, 0_c_int32_t, trim(visc_method) == 'avg_strength')
    p%arlx1i = arlx1i;  p%denom1 = denom1;  p%revp = revp;  p%brlx = brlx
    p%e_factor = e_factor;  p%epp2i = epp2i;  p%capping = capping;  p%Ktens = Ktens
    p%u0 = u0;  p%cosw = cosw;  p%sinw = sinw;  p%rhow = rhow;  p%deltaminEVP = deltaminEVP

    imaskT = merge(1_c_int32_t, 0_c_int32_t, iceTmask);  imaskU = merge(1_c_int32_t, 0_c_int32_t, iceUmask)
    imaskE = merge(1_c_int32_t, 0_c_int32_t, iceEmask);  imaskN = merge(1_c_int32_t, 0_c_int32_t, iceNmask)

    f%uvelE = c_loc(uvelE);  f%vvelE = c_loc(vvelE);  f%uvelN = c_loc(uvelN);  f%vvelN = c_loc(vvelN)
    f%uvel = c_loc(uvel);    f%vvel = c_loc(vvel)
    f%stresspT = c_loc(stresspT);  f%stressmT = c_loc(stressmT);  f%stress12T = c_loc(stress12T);  f%stress12U = c_loc(stress12U)
    f%zetax2T = c_loc(zetax2T);  f%etax2T = c_loc(etax2T);  f%etax2U = c_loc(etax2U);  f%strengthU = c_loc(strengthU)
    f%divergU = c_loc(divergU);  f%tensionU = c_loc(tensionU);  f%shearU = c_loc(shearU);  f%deltaU = c_loc(deltaU)
    f%strintxE = c_loc(strintxE);  f%strintyN = c_loc(strintyN);  f%taubxE = c_loc(taubxE);  f%taubyN = c_loc(taubyN)
    f%strength = c_loc(strength)
    f%cdn_ocnE = c_loc(cdn_ocnE);  f%cdn_ocnN = c_loc(cdn_ocnN);  f%aiE = c_loc(aiE);  f%aiN = c_loc(aiN)
    f%uocnE = c_loc(uocnE);  f%vocnE = c_loc(vocnE);  f%uocnN = c_loc(uocnN);  f%vocnN = c_loc(vocnN)
    f%waterxE = c_loc(waterxE);  f%wateryN = c_loc(wateryN);  f%forcexE = c_loc(forcexE);  f%forceyN = c_loc(forceyN)
    f%emassdti = c_loc(emassdti);  f%nmassdti = c_loc(nmassdti);  f%fmE = c_loc(fmE);  f%fmN = c_loc(fmN)
    f%TbE = c_loc(TbE);  f%TbN = c_loc(TbN);  f%rheofactE = c_loc(rheofactE);  f%rheofactN = c_loc(rheofactN)
    f%iceTmask = c_loc(imaskT);  f%iceUmask = c_loc(imaskU);  f%iceEmask = c_loc(imaskE);  f%iceNmask = c_loc(imaskN)

    call check(evp_b200_run_cgrid(p, f), 'evp_b200_run_cgrid')
  end subroutine dyn_evp_b200_run_cgrid

  !---------------------------------------------------------------------
  subroutine dyn_evp_b200_finalize
    call check(evp_b200_finalize(), 'evp_b200_finalize')
    if (allocated(b_ilo)) deallocate(b_ilo, b_ihi, b_jlo, b_jhi, b_iglob, b_jglob, imaskT, imaskU)
  end subroutine dyn_evp_b200_finalize

end module ice_dyn_evp_b200
