! mistra_kpp_mod.f90 - ISO_C_BINDING interface of libmistra_kpp.so (include/mistra_kpp.h)
! for the Fortran host (boundary B2).  NOT compiled in this repository's CI: the image
! has no Fortran compiler; syntax-reviewed only.  See INTEGRATION.md for the patched
! kpp_driver loop (reference: src/kpp.f90:4310-4470).
module mistra_kpp_mod
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: mistra_kpp_opts, mistra_kpp_default_opts, mistra_kpp_query, mistra_kpp_integrate, &
            mistra_kpp_finalize, MISTRA_KPP_GAS, MISTRA_KPP_AER, MISTRA_KPP_TOT

  integer(c_int), parameter :: MISTRA_KPP_GAS = 0, MISTRA_KPP_AER = 1, MISTRA_KPP_TOT = 2

  ! RPAR/IPAR of Rosenbrock_x (gas.f:786-870); zero selects the reference default
  type, bind(C) :: mistra_kpp_opts
     real(c_double)  :: rtol, atol
     real(c_double)  :: hmin, hmax, hstart
     real(c_double)  :: facmin, facmax, facrej, facsafe
     integer(c_int32_t) :: max_steps, autonomous, f32_literals, reserved
  end type mistra_kpp_opts

  interface
     subroutine mistra_kpp_default_opts(o) bind(C, name="mistra_kpp_default_opts")
       import :: mistra_kpp_opts
       type(mistra_kpp_opts), intent(out) :: o
     end subroutine mistra_kpp_default_opts

     function mistra_kpp_query(mech, nvar, nfix, nreact, lu_nonzero) result(rc) &
          bind(C, name="mistra_kpp_query")
       import :: c_int
       integer(c_int), value :: mech
       integer(c_int), intent(out) :: nvar, nfix, nreact, lu_nonzero
       integer(c_int) :: rc
     end function mistra_kpp_query

     ! rconst(NREACT,ncell), fix(NFIX,ncell), var(NVAR,ncell): column-major Fortran
     ! arrays are exactly the row-major [ncell][...] arrays of the C prototype.
     function mistra_kpp_integrate(mech, ncell, rconst, fix, var, t0, t1, o, ierr, stats, &
          hexit, texit, stream) result(rc) bind(C, name="mistra_kpp_integrate")
       import :: c_int, c_int64_t, c_int32_t, c_double, c_ptr, mistra_kpp_opts
       integer(c_int), value :: mech
       integer(c_int64_t), value :: ncell
       real(c_double), intent(in) :: rconst(*), fix(*)
       real(c_double), intent(inout) :: var(*)
       real(c_double), value :: t0, t1
       type(mistra_kpp_opts), intent(in) :: o
       integer(c_int32_t), intent(out) :: ierr(*), stats(8,*)
       real(c_double), intent(out) :: hexit(*), texit(*)
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_kpp_integrate

     function mistra_kpp_finalize() result(rc) bind(C, name="mistra_kpp_finalize")
       import :: c_int
       integer(c_int) :: rc
     end function mistra_kpp_finalize
  end interface
end module mistra_kpp_mod
