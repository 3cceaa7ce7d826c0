program main
  use m_examples_pprts_dump
  implicit none
  character(len=4096) :: fcase, fresult

  if (command_argument_count() .lt. 2) then
    print *, 'usage: ex_pprts_dump <case file> <result file> [PETSc / TenStream options]'
    stop 1
  end if
  call get_command_argument(1, fcase)
  call get_command_argument(2, fresult)
  call ex_pprts_dump(trim(fcase), trim(fresult))
end program
