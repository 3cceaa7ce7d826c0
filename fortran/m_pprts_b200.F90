!> ISO_C_BINDING shim over libpprts_b200.so (include/pprts_b200.h, layer 2).
!>
!> Host code stays in Fortran: this module keeps the reference's call signatures for the pprts path
!> (src/pprts.F90:192-200 -- init_pprts, set_angles, set_optical_properties, solve_pprts, pprts_get_result,
!> destroy_pprts) and forwards them to the handle-based C ABI.  A host model that today does
!>     use m_pprts, only: init_pprts, set_optical_properties, solve_pprts, pprts_get_result, destroy_pprts
!> switches to `use m_pprts_b200` and replaces `class(t_solver)` by `type(t_solver_b200)`.
!>
!> Not compiled in this repository's CI (the build image has no Fortran compiler); it only uses F2003 features.
module m_pprts_b200
  use iso_c_binding
  implicit none
  private
  public :: t_solver_b200, t_lut_b200, lut_open, lut_close, init_pprts, set_angles, set_optical_properties, &
    & solve_pprts, pprts_get_result, destroy_pprts, t_pprts_buildings_b200, need_new_solution, slope_correction_fluxes, &
    & smooth_surface_fluxes, solve_bands, allreduce_accumulated, get_accumulated

  integer, parameter :: ireals = c_double, iintegers = c_int

  type t_lut_b200
    type(c_ptr) :: h = c_null_ptr
  end type
  type t_solver_b200
    type(c_ptr) :: h = c_null_ptr
    integer(iintegers) :: Nz = 0, Nx = 0, Ny = 0
  end type

  !> the data members of t_pprts_buildings (src/buildings.F90:40-69) that the pprts path reads and fills; a host that
  !> already holds a type(t_pprts_buildings) passes its iface / albedo / planck arrays on unchanged
  type t_pprts_buildings_b200
    integer(c_int), allocatable :: iface(:)   !< 1-based flattened (6,Nz,Nx,Ny) face indices (buildings.F90:215-222)
    real(ireals), allocatable :: albedo(:), planck(:)
    real(ireals), allocatable :: edir(:), incoming(:), outgoing(:)
  end type

  !> mirrors pprts_b200_config (include/pprts_b200.h)
  type, bind(c) :: c_config
    integer(c_int) :: solver_id, Nz, Nx, Ny
    real(c_double) :: dx, dy
    type(c_ptr) :: dz1d, dz3d
    real(c_double) :: sundir(3)
    real(c_double) :: twostr_ratio
    integer(c_int) :: coeff_fp32, by_coeff_divergence
    real(c_double) :: rtol, atol
    integer(c_int) :: maxit_dir, maxit_diff, precond, device, keep_solutions, nxproc, nyproc, xrank, yrank, &
      & diff_explicit, krylov_fp32, collapseindex
    real(c_double) :: max_solution_err, max_solution_time
    integer(c_int) :: geometric_coeffs
    integer(c_int) :: accept_incomplete_solve, gmres_restart, max_cached_solutions, open_bc
  end type

  !> mirrors pprts_b200_bands (include/pprts_b200.h): the band loop of the spectral callers in one call
  type, bind(c) :: c_bands
    integer(c_int) :: nbands, lsolar, uid0, ldelta_scaling
    real(c_double) :: albedo
    type(c_ptr) :: albedo_2d
    type(c_ptr) :: kabs, ksca, g
    integer(c_long) :: stride_kabs, stride_ksca, stride_g
    type(c_ptr) :: planck, planck_srfc
    integer(c_long) :: stride_planck, stride_planck_srfc
    type(c_ptr) :: edirTOA, weights
    integer(c_int) :: accumulate_only
    integer(c_int) :: have_time
    real(c_double) :: time
  end type

  interface
    subroutine c_default_config(cfg) bind(c, name='pprts_b200_default_config')
      import; type(c_config), intent(out) :: cfg
    end subroutine
    integer(c_int) function c_lut_open(solver_id, pT, pS, pD, ndir, axes, lut) bind(c, name='pprts_b200_lut_open')
      import; integer(c_int), value :: solver_id
      character(kind=c_char), intent(in) :: pT(*), pS(*), pD(*)
      integer(c_int), intent(in) :: ndir(6)
      type(c_ptr), intent(in) :: axes(6)
      type(c_ptr), intent(out) :: lut
    end function
    integer(c_int) function c_lut_close(lut) bind(c, name='pprts_b200_lut_close')
      import; type(c_ptr), value :: lut
    end function
    integer(c_int) function c_create(cfg, lut, solver) bind(c, name='pprts_b200_create')
      import; type(c_config), intent(in) :: cfg
      type(c_ptr), value :: lut
      type(c_ptr), intent(out) :: solver
    end function
    integer(c_int) function c_destroy(solver) bind(c, name='pprts_b200_destroy')
      import; type(c_ptr), value :: solver
    end function
    integer(c_int) function c_set_angles(solver, sundir) bind(c, name='pprts_b200_set_angles')
      import; type(c_ptr), value :: solver
      real(c_double), intent(in) :: sundir(3)
    end function
    integer(c_int) function c_set_optprop(solver, albedo, kabs, ksca, g, planck, planck_srfc, albedo_2d, ldelta) &
        & bind(c, name='pprts_b200_set_optical_properties')
      import; type(c_ptr), value :: solver
      real(c_double), value :: albedo
      type(c_ptr), value :: kabs, ksca, g, planck, planck_srfc, albedo_2d
      integer(c_int), value :: ldelta
    end function
    integer(c_int) function c_solve(solver, lthermal, lsolar, edirTOA, uid) bind(c, name='pprts_b200_solve')
      import; type(c_ptr), value :: solver
      integer(c_int), value :: lthermal, lsolar, uid
      real(c_double), value :: edirTOA
    end function
    integer(c_int) function c_get_result(solver, uid, edn, eup, abso, edir) bind(c, name='pprts_b200_get_result')
      import; type(c_ptr), value :: solver
      integer(c_int), value :: uid
      type(c_ptr), value :: edn, eup, abso, edir
    end function
    integer(c_int) function c_nccl_unique_id(id) bind(c, name='pprts_b200_nccl_unique_id')
      import; character(kind=c_char), intent(out) :: id(128)
    end function
    integer(c_int) function c_comm_init(solver, nranks, rank, id) bind(c, name='pprts_b200_comm_init')
      import; type(c_ptr), value :: solver
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id(128)
    end function
    integer(c_int) function c_solve_at_time(solver, lthermal, lsolar, edirTOA, uid, time) bind(c, name='pprts_b200_solve_at_time')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: solver
      integer(c_int), value :: lthermal, lsolar, uid
      real(c_double), value :: edirTOA, time
    end function
    integer(c_int) function c_need_new_solution(solver, uid, time, have_time, need) bind(c, name='pprts_b200_need_new_solution')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: solver
      integer(c_int), value :: uid, have_time
      real(c_double), value :: time
      integer(c_int), intent(out) :: need
    end function
    integer(c_int) function c_slope_correction(solver, lslope, latm, edir) bind(c, name='pprts_b200_slope_correction_fluxes')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver, edir
      integer(c_int), value :: lslope, latm
    end function
    integer(c_int) function c_smooth_surface_fluxes(solver, radius, edn, eup) bind(c, name='pprts_b200_smooth_surface_fluxes')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: solver, edn, eup
      real(c_double), value :: radius
    end function
    integer(c_int) function c_set_buildings(solver, nfaces, iface, albedo, planck) bind(c, name='pprts_b200_set_buildings')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver, iface, albedo, planck
      integer(c_int), value :: nfaces
    end function
    integer(c_int) function c_get_buildings_result(solver, uid, edir, incoming, outgoing) &
        & bind(c, name='pprts_b200_get_buildings_result')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver, edir, incoming, outgoing
      integer(c_int), value :: uid
    end function
    integer(c_int) function c_set_abso_in_buildings(solver, uid, val) bind(c, name='pprts_b200_set_abso_in_buildings')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: solver
      integer(c_int), value :: uid
      real(c_double), value :: val
    end function
    ! overlapped host transfers for a caller's own band loop (INTEGRATION.md section 3)
    integer(c_int) function c_stage_optical_properties_async(solver, slot, albedo, kabs, ksca, g, planck, planck_srfc, albedo_2d) &
        & bind(c, name='pprts_b200_stage_optical_properties_async')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: solver
      integer(c_int), value :: slot
      real(c_double), value :: albedo
      type(c_ptr), value :: kabs, ksca, g, planck, planck_srfc, albedo_2d   ! c_loc of the arrays, c_null_ptr if absent
    end function
    integer(c_int) function c_set_optical_properties_staged(solver, slot, ldelta_scaling) &
        & bind(c, name='pprts_b200_set_optical_properties_staged')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver
      integer(c_int), value :: slot, ldelta_scaling
    end function
    integer(c_int) function c_get_result_async(solver, uid, edn, eup, abso, edir) bind(c, name='pprts_b200_get_result_async')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver
      integer(c_int), value :: uid
      type(c_ptr), value :: edn, eup, abso, edir
    end function
    integer(c_int) function c_wait_results(solver) bind(c, name='pprts_b200_wait_results')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver
    end function
    integer(c_int) function c_solve_bands(solver, bands, edn, eup, abso, edir, niter_diff) bind(c, name='pprts_b200_solve_bands')
      import :: c_ptr, c_int, c_bands
      type(c_ptr), value :: solver
      type(c_bands), intent(in) :: bands
      type(c_ptr), value :: edn, eup, abso, edir, niter_diff
    end function
    integer(c_int) function c_allreduce_accumulated(solver) bind(c, name='pprts_b200_allreduce_accumulated')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver
    end function
    integer(c_int) function c_get_accumulated(solver, edn, eup, abso, edir) bind(c, name='pprts_b200_get_accumulated')
      import :: c_ptr, c_int
      type(c_ptr), value :: solver, edn, eup, abso, edir
    end function
    function c_last_error() bind(c, name='pprts_b200_last_error') result(msg)
      import; type(c_ptr) :: msg
    end function
  end interface

contains

  !> CHKERR (src/helper_functions.fypp:888-906): print and stop; a host model may replace this by mpi_abort
  subroutine chkerr(ierr, where)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: cmsg(:)
    integer :: n
    if (ierr .eq. 0) return
    call c_f_pointer(c_last_error(), cmsg, [1024])
    n = 1
    do while (n .lt. 1024 .and. cmsg(n) .ne. c_null_char); n = n + 1; end do
    print *, 'pprts_b200 :: '//where//' failed (', ierr, '): ', cmsg(1:n - 1)
    error stop 1
  end subroutine

  subroutine lut_open(lut, solver_id, fname_Tdir, fname_Sdir, fname_Sdiff, tau, w0, aspect_zx, g, phi, theta)
    type(t_lut_b200), intent(out) :: lut
    integer(iintegers), intent(in) :: solver_id
    character(len=*), intent(in) :: fname_Tdir, fname_Sdir, fname_Sdiff
    real(c_float), intent(in), target :: tau(:), w0(:), aspect_zx(:), g(:), phi(:), theta(:)
    integer(c_int) :: ndir(6)
    type(c_ptr) :: axes(6)
    ndir = [size(tau), size(w0), size(aspect_zx), size(g), size(phi), size(theta)]
    axes = [c_loc(tau), c_loc(w0), c_loc(aspect_zx), c_loc(g), c_loc(phi), c_loc(theta)]
    call chkerr(c_lut_open(solver_id, trim(fname_Tdir)//c_null_char, trim(fname_Sdir)//c_null_char, &
      & trim(fname_Sdiff)//c_null_char, ndir, axes, lut%h), 'lut_open')
  end subroutine

  subroutine lut_close(lut)
    type(t_lut_b200), intent(inout) :: lut
    call chkerr(c_lut_close(lut%h), 'lut_close'); lut%h = c_null_ptr
  end subroutine

  !> init_pprts (src/pprts.F90:213).  One MPI rank drives one GPU.  With nxproc/nyproc (columns per process along x/y,
  !> exactly the reference's optional arguments) the domain is decomposed DMDA-style: Nx, Ny are then LOCAL sizes, the
  !> process grid is x-fastest (rank = xi + yi*size(nxproc), src/pprts_base.F90:745-773), and the NCCL communicator of
  !> the halo exchange is bootstrapped over icomm (rank 0 creates the NCCL id, MPI_Bcast distributes it).
  !> The trailing optional arguments carry what the reference reads from its options database:
  !> collapseindex (init_pprts argument), -pprts_geometric_coeffs, -max_solution_err, -max_solution_time,
  !> -twostr_ratio, -solar_diff_ksp_rtol / -thermal_diff_ksp_rtol (one value for both here), -accept_incomplete_solve,
  !> -pprts_open_bc
  subroutine init_pprts(icomm, Nz, Nx, Ny, dx, dy, sundir, solver, lut, solver_id, dz1d, dz3d, nxproc, nyproc, device, &
      & collapseindex, lgeometric_coeffs, max_solution_err, max_solution_time, twostr_ratio, ksp_rtol, &
      & ltop_bottom_faces_planar, ltop_bottom_planes_parallel, laccept_incomplete_solve, lopen_bc)
    use mpi
    integer, intent(in) :: icomm
    integer(iintegers), intent(in) :: Nz, Nx, Ny
    real(ireals), intent(in) :: dx, dy, sundir(3)
    type(t_solver_b200), intent(inout) :: solver
    type(t_lut_b200), intent(in) :: lut
    integer(iintegers), intent(in) :: solver_id
    real(ireals), intent(in), target, optional :: dz1d(:), dz3d(:, :, :)
    integer(iintegers), intent(in), optional :: nxproc(:), nyproc(:)
    integer(iintegers), intent(in), optional :: device
    integer(iintegers), intent(in), optional :: collapseindex
    logical, intent(in), optional :: lgeometric_coeffs, ltop_bottom_faces_planar, ltop_bottom_planes_parallel
    logical, intent(in), optional :: laccept_incomplete_solve  !< -accept_incomplete_solve (src/pprts.F90:4270-4275)
    logical, intent(in), optional :: lopen_bc  !< -pprts_open_bc (src/pprts.F90:462-463, src/pprts_base.F90:1189-1316)
    real(ireals), intent(in), optional :: max_solution_err, max_solution_time, twostr_ratio, ksp_rtol
    type(c_config) :: cfg
    character(kind=c_char) :: nccl_id(128)
    integer :: myid, nranks, ierr
    call c_default_config(cfg)
    cfg%solver_id = solver_id; cfg%Nz = Nz; cfg%Nx = Nx; cfg%Ny = Ny; cfg%dx = dx; cfg%dy = dy; cfg%sundir = sundir
    if (present(dz1d)) cfg%dz1d = c_loc(dz1d)
    if (present(dz3d)) cfg%dz3d = c_loc(dz3d)
    if (present(device)) cfg%device = device
    if (present(collapseindex)) cfg%collapseindex = collapseindex
    if (present(lgeometric_coeffs)) cfg%geometric_coeffs = merge(1, 0, lgeometric_coeffs)
    if (cfg%geometric_coeffs .ne. 0) then ! -pprts_top_bottom_faces_planar no / -pprts_top_bottom_planes_parallel no
      if (present(ltop_bottom_faces_planar)) &
        & cfg%geometric_coeffs = cfg%geometric_coeffs + merge(0, 2, ltop_bottom_faces_planar)
      if (present(ltop_bottom_planes_parallel)) &
        & cfg%geometric_coeffs = cfg%geometric_coeffs + merge(0, 4, ltop_bottom_planes_parallel)
    end if
    if (present(max_solution_err)) cfg%max_solution_err = max_solution_err
    if (present(max_solution_time)) cfg%max_solution_time = max_solution_time
    if (present(twostr_ratio)) cfg%twostr_ratio = twostr_ratio
    if (present(ksp_rtol)) cfg%rtol = ksp_rtol
    if (present(laccept_incomplete_solve)) cfg%accept_incomplete_solve = merge(1, 0, laccept_incomplete_solve)
    if (present(lopen_bc)) cfg%open_bc = merge(1, 0, lopen_bc)
    if (present(nxproc) .and. present(nyproc)) then
      call mpi_comm_rank(icomm, myid, ierr)
      call mpi_comm_size(icomm, nranks, ierr)
      cfg%nxproc = size(nxproc); cfg%nyproc = size(nyproc)
      cfg%xrank = modulo(myid, cfg%nxproc); cfg%yrank = myid / cfg%nxproc
    end if
    call chkerr(c_create(cfg, lut%h, solver%h), 'init_pprts')
    solver%Nz = Nz - max(cfg%collapseindex, 1) + 1  ! results live on the collapsed solver grid (src/pprts.F90:692-706)
    solver%Nx = Nx; solver%Ny = Ny
    if (cfg%nxproc * cfg%nyproc .gt. 1) then
      if (myid .eq. 0) call chkerr(c_nccl_unique_id(nccl_id), 'nccl_unique_id')
      call mpi_bcast(nccl_id, 128, MPI_CHARACTER, 0, icomm, ierr)
      call chkerr(c_comm_init(solver%h, int(nranks, c_int), int(myid, c_int), nccl_id), 'comm_init')
    end if
  end subroutine

  subroutine set_angles(solver, sundir)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), intent(in) :: sundir(3)
    call chkerr(c_set_angles(solver%h, sundir), 'set_angles')
  end subroutine

  !> set_optical_properties (src/pprts.F90:1764): local arrays (Nz,Nx,Ny); planck (Nz+1,Nx,Ny)
  subroutine set_optical_properties(solver, albedo, kabs, ksca, g, planck, planck_srfc, albedo_2d, ldelta_scaling)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), intent(in) :: albedo
    real(ireals), intent(in), target, contiguous, optional :: kabs(:, :, :), ksca(:, :, :), g(:, :, :), planck(:, :, :)
    real(ireals), intent(in), target, contiguous, optional :: planck_srfc(:, :), albedo_2d(:, :)
    logical, intent(in), optional :: ldelta_scaling
    type(c_ptr) :: p(6)
    integer(c_int) :: ldelta
    p = c_null_ptr
    if (present(kabs)) p(1) = c_loc(kabs)
    if (present(ksca)) p(2) = c_loc(ksca)
    if (present(g)) p(3) = c_loc(g)
    if (present(planck)) p(4) = c_loc(planck)
    if (present(planck_srfc)) p(5) = c_loc(planck_srfc)
    if (present(albedo_2d)) p(6) = c_loc(albedo_2d)
    ldelta = 1
    if (present(ldelta_scaling)) ldelta = merge(1, 0, ldelta_scaling)
    call chkerr(c_set_optprop(solver%h, albedo, p(1), p(2), p(3), p(4), p(5), p(6), ldelta), 'set_optical_properties')
  end subroutine

  !> solve_pprts (src/pprts.F90:2487)
  subroutine solve_pprts(solver, lthermal, lsolar, edirTOA, opt_solution_uid, opt_solution_time, opt_buildings)
    type(t_solver_b200), intent(inout) :: solver
    logical, intent(in) :: lthermal, lsolar
    real(ireals), intent(in) :: edirTOA
    integer(iintegers), intent(in), optional :: opt_solution_uid
    real(ireals), intent(in), optional :: opt_solution_time
    type(t_pprts_buildings_b200), intent(in), target, optional :: opt_buildings
    integer(c_int) :: uid
    type(c_ptr) :: ppl
    uid = 0; if (present(opt_solution_uid)) uid = opt_solution_uid
    if (present(opt_buildings)) then  ! src/pprts.F90:2494; the list stays attached until the next call without it
      ppl = c_null_ptr
      if (allocated(opt_buildings%planck)) ppl = c_loc(opt_buildings%planck)
      call chkerr(c_set_buildings(solver%h, size(opt_buildings%iface), c_loc(opt_buildings%iface), &
        & c_loc(opt_buildings%albedo), ppl), 'solve_pprts(opt_buildings)')
    else
      call chkerr(c_set_buildings(solver%h, 0_c_int, c_null_ptr, c_null_ptr, c_null_ptr), 'solve_pprts')
    end if
    if (present(opt_solution_time)) then
      call chkerr(c_solve_at_time(solver%h, merge(1, 0, lthermal), merge(1, 0, lsolar), edirTOA, uid, opt_solution_time), &
        & 'solve_pprts')
    else
      call chkerr(c_solve(solver%h, merge(1, 0, lthermal), merge(1, 0, lsolar), edirTOA, uid), 'solve_pprts')
    end if
  end subroutine

  !> need_new_solution (src/adaptive_spectral_integration.F90:38): the band loops' test before every solve
  !> (rrtmg/rrtmg/pprts_rrtmg.F90:1001); the solution is addressed by its uid instead of solver%solutions(uid)
  logical function need_new_solution(solver, uid, time)
    type(t_solver_b200), intent(in) :: solver
    integer(iintegers), intent(in) :: uid
    real(ireals), intent(in), optional :: time
    integer(c_int) :: need
    if (present(time)) then
      call chkerr(c_need_new_solution(solver%h, uid, time, 1_c_int, need), 'need_new_solution')
    else
      call chkerr(c_need_new_solution(solver%h, uid, 0._c_double, 0_c_int, need), 'need_new_solution')
    end if
    need_new_solution = need .ne. 0
  end function

  !> pprts_get_result (src/pprts.F90:5799): allocates the outputs if needed, like the reference
  !> set_abso_in_buildings: the value of -pprts_set_abso_in_buildings (src/pprts.F90:5986-6009), an argument here
  subroutine pprts_get_result(solver, redn, reup, rabso, redir, opt_solution_uid, opt_buildings, set_abso_in_buildings)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), allocatable, target, intent(inout) :: redn(:, :, :), reup(:, :, :), rabso(:, :, :)
    real(ireals), allocatable, target, intent(inout), optional :: redir(:, :, :)
    integer(iintegers), intent(in), optional :: opt_solution_uid
    type(t_pprts_buildings_b200), intent(inout), target, optional :: opt_buildings
    real(ireals), intent(in), optional :: set_abso_in_buildings
    integer(c_int) :: uid
    type(c_ptr) :: pdir
    uid = 0; if (present(opt_solution_uid)) uid = opt_solution_uid
    if (.not. allocated(redn)) allocate (redn(solver%Nz + 1, solver%Nx, solver%Ny))
    if (.not. allocated(reup)) allocate (reup(solver%Nz + 1, solver%Nx, solver%Ny))
    if (.not. allocated(rabso)) allocate (rabso(solver%Nz, solver%Nx, solver%Ny))
    pdir = c_null_ptr
    if (present(redir)) then
      if (.not. allocated(redir)) allocate (redir(solver%Nz + 1, solver%Nx, solver%Ny))
      pdir = c_loc(redir)
    end if
    call chkerr(c_get_result(solver%h, uid, c_loc(redn), c_loc(reup), c_loc(rabso), pdir), 'pprts_get_result')
    if (present(opt_buildings)) then  ! fill_buildings_arr, src/pprts.F90:5906-5910, :6011-6160
      if (.not. allocated(opt_buildings%edir)) allocate (opt_buildings%edir(size(opt_buildings%iface)))
      if (.not. allocated(opt_buildings%incoming)) allocate (opt_buildings%incoming(size(opt_buildings%iface)))
      if (.not. allocated(opt_buildings%outgoing)) allocate (opt_buildings%outgoing(size(opt_buildings%iface)))
      call chkerr(c_get_buildings_result(solver%h, uid, c_loc(opt_buildings%edir), c_loc(opt_buildings%incoming), &
        & c_loc(opt_buildings%outgoing)), 'pprts_get_result(opt_buildings)')
      if (present(set_abso_in_buildings)) &
        & call chkerr(c_set_abso_in_buildings(solver%h, uid, real(set_abso_in_buildings, c_double)), 'set_abso_in_buildings')
    end if
  end subroutine

  !> slope_correction_fluxes (src/pprts_postprocess.F90:130); the options -pprts_slope_correction / -pprts_atm_correction
  !> are arguments here
  subroutine slope_correction_fluxes(solver, edir, lslope_correction, latm_correction)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), allocatable, target, intent(inout) :: edir(:, :, :)
    logical, intent(in) :: lslope_correction, latm_correction
    call chkerr(c_slope_correction(solver%h, merge(1, 0, lslope_correction), merge(1, 0, latm_correction), c_loc(edir)), &
      & 'slope_correction_fluxes')
  end subroutine

  !> smooth_surface_fluxes (src/pprts_postprocess.F90:44); the value of -pprts_smooth_srfc_flx is the argument radius
  !> (absent = option not set: nothing happens, as in the reference)
  subroutine smooth_surface_fluxes(solver, edn, eup, radius)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), allocatable, target, intent(inout) :: edn(:, :, :), eup(:, :, :)
    real(ireals), intent(in), optional :: radius
    if (.not. present(radius)) return
    call chkerr(c_smooth_surface_fluxes(solver%h, real(radius, c_double), c_loc(edn), c_loc(eup)), 'smooth_surface_fluxes')
  end subroutine

  !> The band loop of the spectral callers (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053 solar, :621-708 thermal) in one call:
  !> set_optical_properties -> solve_pprts -> pprts_get_result -> weighted sum for every band, the upload of band ib+1
  !> overlapping the solve of band ib, the spectral sum kept on the device.  kabs / ksca / g: (Nz,Nx,Ny,nbands);
  !> planck (Nz+1,Nx,Ny,nbands), planck_srfc (Nx,Ny,nbands) for thermal bands; edirTOA(nbands) for solar bands; weights
  !> default to 1; uids are uid0 + ib - 1 (default: the reference's 1.. solar, 501.. thermal).  With opt_time, bands whose
  !> stored solution is still good (need_new_solution) are not solved again (niter = -1).  The sums come back in
  !> edn / eup / abso / edir if present; without them they stay in the device accumulator (bands sharded over several
  !> processes: allreduce_accumulated, then get_accumulated).
  subroutine solve_bands(solver, lsolar, albedo, kabs, ksca, g, planck, planck_srfc, edirTOA, weights, uid0, opt_time, &
      & accumulate_only, edn, eup, abso, edir, niter)
    type(t_solver_b200), intent(inout) :: solver
    logical, intent(in) :: lsolar
    real(ireals), intent(in) :: albedo
    real(ireals), intent(in), contiguous, target :: kabs(:, :, :, :), ksca(:, :, :, :), g(:, :, :, :)
    real(ireals), intent(in), contiguous, target, optional :: planck(:, :, :, :), planck_srfc(:, :, :)
    real(ireals), intent(in), contiguous, target, optional :: edirTOA(:), weights(:)
    integer(iintegers), intent(in), optional :: uid0
    real(ireals), intent(in), optional :: opt_time
    logical, intent(in), optional :: accumulate_only
    real(ireals), allocatable, target, intent(inout), optional :: edn(:, :, :), eup(:, :, :), abso(:, :, :), edir(:, :, :)
    integer(c_int), intent(out), contiguous, target, optional :: niter(:)
    type(c_bands) :: b
    type(c_ptr) :: pdn, pup, pab, pdi, pni

    b%nbands = size(kabs, 4)
    if (b%nbands .eq. 0) return  ! a process without bands of this kind (more processes than bands)
    b%lsolar = merge(1, 0, lsolar)
    b%uid0 = merge(1, 501, lsolar); if (present(uid0)) b%uid0 = uid0
    b%ldelta_scaling = 1
    b%albedo = albedo
    b%albedo_2d = c_null_ptr
    b%kabs = c_loc(kabs); b%ksca = c_loc(ksca); b%g = c_loc(g)
    b%stride_kabs = size(kabs, kind=c_long) / b%nbands
    b%stride_ksca = size(ksca, kind=c_long) / b%nbands
    b%stride_g = size(g, kind=c_long) / b%nbands
    b%planck = c_null_ptr; b%stride_planck = 0
    if (present(planck)) then
      b%planck = c_loc(planck); b%stride_planck = size(planck, kind=c_long) / b%nbands
    end if
    b%planck_srfc = c_null_ptr; b%stride_planck_srfc = 0
    if (present(planck_srfc)) then
      b%planck_srfc = c_loc(planck_srfc); b%stride_planck_srfc = size(planck_srfc, kind=c_long) / b%nbands
    end if
    b%edirTOA = c_null_ptr; if (present(edirTOA)) b%edirTOA = c_loc(edirTOA)
    b%weights = c_null_ptr; if (present(weights)) b%weights = c_loc(weights)
    b%accumulate_only = 0; if (present(accumulate_only)) b%accumulate_only = merge(1, 0, accumulate_only)
    b%have_time = 0; b%time = 0
    if (present(opt_time)) then
      b%have_time = 1; b%time = opt_time
    end if

    pdn = c_null_ptr; pup = c_null_ptr; pab = c_null_ptr; pdi = c_null_ptr; pni = c_null_ptr
    if (present(edn)) then
      if (.not. allocated(edn)) allocate (edn(solver%Nz + 1, solver%Nx, solver%Ny))
      pdn = c_loc(edn)
    end if
    if (present(eup)) then
      if (.not. allocated(eup)) allocate (eup(solver%Nz + 1, solver%Nx, solver%Ny))
      pup = c_loc(eup)
    end if
    if (present(abso)) then
      if (.not. allocated(abso)) allocate (abso(solver%Nz, solver%Nx, solver%Ny))
      pab = c_loc(abso)
    end if
    if (present(edir)) then
      if (.not. allocated(edir)) allocate (edir(solver%Nz + 1, solver%Nx, solver%Ny))
      pdi = c_loc(edir)
    end if
    if (present(niter)) pni = c_loc(niter)
    call chkerr(c_solve_bands(solver%h, b, pdn, pup, pab, pdi, pni), 'solve_bands')
  end subroutine

  !> sum of the device accumulators over all processes (the final flux reduction of band sharding); needs comm_init
  subroutine allreduce_accumulated(solver)
    type(t_solver_b200), intent(inout) :: solver
    call chkerr(c_allreduce_accumulated(solver%h), 'allreduce_accumulated')
  end subroutine

  !> the spectral sums held in the device accumulator -> host arrays (allocated if needed)
  subroutine get_accumulated(solver, edn, eup, abso, edir)
    type(t_solver_b200), intent(inout) :: solver
    real(ireals), allocatable, target, intent(inout) :: edn(:, :, :), eup(:, :, :), abso(:, :, :), edir(:, :, :)
    if (.not. allocated(edn)) allocate (edn(solver%Nz + 1, solver%Nx, solver%Ny))
    if (.not. allocated(eup)) allocate (eup(solver%Nz + 1, solver%Nx, solver%Ny))
    if (.not. allocated(abso)) allocate (abso(solver%Nz, solver%Nx, solver%Ny))
    if (.not. allocated(edir)) allocate (edir(solver%Nz + 1, solver%Nx, solver%Ny))
    call chkerr(c_get_accumulated(solver%h, c_loc(edn), c_loc(eup), c_loc(abso), c_loc(edir)), 'get_accumulated')
  end subroutine

  subroutine destroy_pprts(solver, lfinalizepetsc)
    type(t_solver_b200), intent(inout) :: solver
    logical, intent(in), optional :: lfinalizepetsc
    call chkerr(c_destroy(solver%h), 'destroy_pprts'); solver%h = c_null_ptr
    if (.false. .and. present(lfinalizepetsc)) print *, lfinalizepetsc
  end subroutine
end module
