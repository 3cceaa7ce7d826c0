!> Runs ONE pprts solve of the real reference (TenStream) on the inputs of a case file and writes the fluxes to a result
!> file, so that a machine that has the reference's toolchain (Fortran + MPI + PETSc + NetCDF + fypp) and its LUT files can
!> produce reference numbers that tests/test_gpu_vs_reference_dump.py holds this repo's CUDA path to.
!> File formats: tests/tools/casefile.py (raw little-endian streams).  See tools/reference_dump/README.md for the build
!> (drop this file and ex_pprts_dump.F90 into the reference's examples/pprts/, its CMake globs them) and the run line.
!> Written against the reference's public API only (src/pprts.F90:192-200); never compiled in this repo's container.
module m_examples_pprts_dump
  use iso_fortran_env, only: int32, real64
  use m_data_parameters, only: init_mpi_data_parameters, finalize_mpi, iintegers, ireals, mpiint
  use m_pprts, only: init_pprts, set_optical_properties, solve_pprts, pprts_get_result
  use m_pprts_base, only: t_solver, allocate_pprts_solver_from_commandline, destroy_pprts
  use m_helper_functions, only: CHKERR
  use mpi, only: MPI_COMM_WORLD
  implicit none

  integer(int32), parameter :: CASE_MAGIC = int(z'31435050', int32), RESULT_MAGIC = int(z'31525050', int32)

contains
  subroutine ex_pprts_dump(fcase, fresult)
    character(len=*), intent(in) :: fcase, fresult
    integer(int32) :: head(8)
    real(real64) :: scal(9)
    real(real64), allocatable :: dz64(:), a64(:, :, :)
    real(ireals), allocatable, dimension(:) :: dz1d
    real(ireals), allocatable, dimension(:, :, :) :: kabs, ksca, g, plck
    real(ireals), allocatable, dimension(:, :, :) :: fdir, fdn, fup, fdiv
    class(t_solver), allocatable :: solver
    character(len=8) :: solvername
    integer(iintegers) :: Nz, Nx, Ny
    integer(mpiint) :: ierr, numnodes
    integer :: u

    call init_mpi_data_parameters(MPI_COMM_WORLD)
    call mpi_comm_size(MPI_COMM_WORLD, numnodes, ierr)
    if (numnodes .ne. 1) call CHKERR(1_mpiint, 'pprts_dump runs on ONE rank: the case file holds the global domain')

    open (newunit=u, file=trim(fcase), access='stream', form='unformatted', status='old', action='read')
    read (u) head
    if (head(1) .ne. CASE_MAGIC) call CHKERR(1_mpiint, 'not a pprts case file: '//trim(fcase))
    Nz = head(2); Nx = head(3); Ny = head(4)
    read (u) scal  ! dx, dy, sundir(3), albedo, edirTOA, rtol, atol (tolerances: pass them as PETSc options, see README)
    allocate (dz64(Nz), a64(Nz, Nx, Ny))
    read (u) dz64
    allocate (dz1d(Nz), kabs(Nz, Nx, Ny), ksca(Nz, Nx, Ny), g(Nz, Nx, Ny))
    dz1d = real(dz64, ireals)
    read (u) a64; kabs = real(a64, ireals)
    read (u) a64; ksca = real(a64, ireals)
    read (u) a64; g = real(a64, ireals)
    deallocate (a64)
    if (head(8) .ne. 0) then
      allocate (a64(Nz + 1, Nx, Ny), plck(Nz + 1, Nx, Ny))
      read (u) a64; plck = real(a64, ireals)
      deallocate (a64)
    end if
    close (u)

    select case (head(5))
    case (310); solvername = '3_10'
    case (810); solvername = '8_10'
    case (324); solvername = '3_24'
    case default
      call CHKERR(1_mpiint, 'case file asks for a solver this repo does not cover')
    end select
    call allocate_pprts_solver_from_commandline(solver, trim(solvername), ierr); call CHKERR(ierr)

    call init_pprts(MPI_COMM_WORLD, Nz, Nx, Ny, real(scal(1), ireals), real(scal(2), ireals), &
      & real(scal(3:5), ireals), solver, dz1d=dz1d)
    if (allocated(plck)) then
      call set_optical_properties(solver, real(scal(6), ireals), kabs, ksca, g, plck)
    else
      call set_optical_properties(solver, real(scal(6), ireals), kabs, ksca, g)
    end if
    call solve_pprts(solver, lthermal=head(7) .ne. 0, lsolar=head(6) .ne. 0, edirTOA=real(scal(7), ireals))

    allocate (fdir(Nz + 1, Nx, Ny), fdn(Nz + 1, Nx, Ny), fup(Nz + 1, Nx, Ny), fdiv(Nz, Nx, Ny))
    call pprts_get_result(solver, fdn, fup, fdiv, fdir)

    open (newunit=u, file=trim(fresult), access='stream', form='unformatted', status='replace', action='write')
    write (u) RESULT_MAGIC, int(Nz, int32), int(Nx, int32), int(Ny, int32)
    write (u) real(fdn, real64)
    write (u) real(fup, real64)
    write (u) real(fdir, real64)
    write (u) real(fdiv, real64)
    close (u)

    call destroy_pprts(solver, .true.)
    call finalize_mpi(MPI_COMM_WORLD, .true., .true.)
  end subroutine
end module
