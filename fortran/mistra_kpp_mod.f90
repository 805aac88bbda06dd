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

! mistra_bins_mod / mistra_kon_mod - ISO_C_BINDING interfaces of include/mistra_bins.h and
! include/mistra_kon.h (2-D particle-grid kernels).  The grid structs hold C pointers to the
! host arrays of COMMON /cb50/, /blck06/, /cb44/, /cb49/ (c_loc of the Fortran arrays).
module mistra_bins_mod
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: mistra_bins_grid
     integer(c_int32_t) :: nka, nkt, ka, nkc_l, ial_first, reserved
     type(c_ptr) :: kw, en, rq          ! kw(nka) int32, en(nka), rq(nkt,nka)
  end type mistra_bins_grid
  interface
     ! ff(nkt,nka,ncell), cm(nkc,ncell), sion1(j6,nkc,ncell) -> sap(nkc,ncell), smp(nkc,ncell), sion1o(9,nkc,ncell)
     function mistra_bins_snapshot(g, ncell, ff, cm, sion1, sap, smp, sion1o, stream) result(rc) &
          bind(C, name="mistra_bins_snapshot")
       import :: c_int, c_int64_t, c_double, c_ptr, mistra_bins_grid
       type(mistra_bins_grid), intent(in) :: g
       integer(c_int64_t), value :: ncell
       real(c_double), intent(in) :: ff(*), cm(*), sion1(*)
       real(c_double), intent(out) :: sap(*), smp(*)
       real(c_double), intent(inout) :: sion1o(*)
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_bins_snapshot
     function mistra_bins_redistribute(g, ncell, ff, cm, cw, sap, smp, sion1o, sion1, sl1, nwarn, stream) &
          result(rc) bind(C, name="mistra_bins_redistribute")
       import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr, mistra_bins_grid
       type(mistra_bins_grid), intent(in) :: g
       integer(c_int64_t), value :: ncell
       real(c_double), intent(inout) :: ff(*), sion1(*), sl1(*)
       real(c_double), intent(in) :: cm(*), cw(*), sap(*), smp(*), sion1o(*)
       integer(c_int32_t), intent(out) :: nwarn(*)
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_bins_redistribute
  end interface
end module mistra_bins_mod

module mistra_kon_mod
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: mistra_kon_grid
     integer(c_int32_t) :: nka, nkt
     real(c_double) :: a0m, dlne
     type(c_ptr) :: en, rn, b0m, ew, e, dew, rw, qabs   ! rw(nkt,nka), qabs(18,nkt,nka,3)
     type(c_ptr) :: kw, rq                              ! kw(nka) int32, rq(nkt,nka)  (mistra_kon_layers)
     integer(c_int32_t) :: ka, reserved
  end type mistra_kon_grid
  ! mistra_kon_state: c_loc pointers to the COMMON arrays of kon, first layer of the batch
  type, bind(C) :: mistra_kon_state
     type(c_ptr) :: ff, t, talt, xm1, xm1a, feu, dfddt, xm2, dtcon, p, totrad, nar
     type(c_ptr) :: vol1_a, vol1_d, part_o_a, part_o_d, part_n_a, part_n_d, vol2, pntot, status
  end type mistra_kon_state
  interface
     ! replaces the layer loop of SUBROUTINE kon (str.f90:4615-4772)
     function mistra_kon_layers(g, ncell, dt, chem, st, stream) result(rc) bind(C, name="mistra_kon_layers")
       import :: c_int, c_int64_t, c_double, c_ptr, mistra_kon_grid, mistra_kon_state
       type(mistra_kon_grid), intent(in) :: g
       integer(c_int64_t), value :: ncell
       real(c_double), value :: dt
       integer(c_int), value :: chem
       type(mistra_kon_state), intent(in) :: st
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_kon_layers
     ! replaces "call subkon(dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr)" (str.f90:4705)
     ! for ncell layers: ffk(nkt,nka,ncell), totr(18,ncell), the scalars as arrays (ncell)
     function mistra_kon_subkon(g, ncell, dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr, &
          status, stream) result(rc) bind(C, name="mistra_kon_subkon")
       import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr, mistra_kon_grid
       type(mistra_kon_grid), intent(in) :: g
       integer(c_int64_t), value :: ncell
       real(c_double), value :: dt
       real(c_double), intent(inout) :: ffk(*), to(*), xm1o(*)
       real(c_double), intent(in) :: totr(*), dfdt(*), feualt(*), pp(*), tn(*), xm1n(*)
       integer(c_int32_t), intent(in) :: kr(*)
       integer(c_int32_t), intent(out) :: status(*)
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_kon_subkon
  end interface
end module mistra_kon_mod

! mistra_konc_mod - ISO_C_BINDING interface of include/mistra_konc.h: SUBROUTINE konc
! (kpp.f90:3370-3585) for all layers 2..nf at once, on the COMMON arrays of /blck07/, /blck08/
! and /blck17/ in place.
module mistra_konc_mod
  use, intrinsic :: iso_c_binding
  implicit none
  ! c_loc pointers to the first layer of the batch: vol1_a(nka,k) ... pntot(nkc,k), sl1(j2,nkc,k),
  ! sion1(j6,nkc,k); warn(3,ncell) int32 or c_null_ptr
  type, bind(C) :: mistra_konc_args
     integer(c_int32_t) :: nka, ka, j2, j6
     type(c_ptr) :: vol1_a, vol1_d, part_o_a, part_o_d, part_n_a, part_n_d, vol2, pntot, sl1, sion1, warn
  end type mistra_konc_args
  interface
     function mistra_konc(ncell, a, stream) result(rc) bind(C, name="mistra_konc")
       import :: c_int, c_int64_t, c_ptr, mistra_konc_args
       integer(c_int64_t), value :: ncell
       type(mistra_konc_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_konc
  end interface
end module mistra_konc_mod

! mistra_cwrc_mod / mistra_fastkmt_mod - ISO_C_BINDING interfaces of include/mistra_cwrc.h and
! include/mistra_fastkmt.h: SUBROUTINE cw_rc (kpp.f90:2152-2414) and SUBROUTINE fast_k_mt_a /
! fast_k_mt_t (kpp.f90:2683-2947, 2421-2676) for all layers at once, on the COMMON arrays in place.
module mistra_cwrc_mod
  use, intrinsic :: iso_c_binding
  implicit none
  ! cloud: int32 copy of the LOGICAL cloudt(nkc,k) of /kpp_l1/ (0 = .false.)
  type, bind(C) :: mistra_cwrc_args
     integer(c_int32_t) :: nka, nkt, ka, ial
     real(c_double) :: xcryssulf, xcrysss, xdelisulf, xdeliss
     type(c_ptr) :: kw, e, rq, ff, feu, cloud, rc, cw, cm, conv2
  end type mistra_cwrc_args
  interface
     function mistra_cwrc(ncell, a, stream) result(rc) bind(C, name="mistra_cwrc")
       import :: c_int, c_int64_t, c_ptr, mistra_cwrc_args
       integer(c_int64_t), value :: ncell
       type(mistra_cwrc_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_cwrc
  end interface
end module mistra_cwrc_mod

module mistra_fastkmt_mod
  use, intrinsic :: iso_c_binding
  implicit none
  ! lex: int32 copy of the DATA lex(nx) table; xkmt(NSPEC,nkc,k) and vt(nkc,k) are updated in place
  type, bind(C) :: mistra_fastkmt_args
     integer(c_int32_t) :: nka, nkt, ka, ial, nkc, nkc_l, nspec, nx
     type(c_ptr) :: lex, kw, rq, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt
  end type mistra_fastkmt_args
  interface
     function mistra_fastkmt(ncell, a, stream) result(rc) bind(C, name="mistra_fastkmt")
       import :: c_int, c_int64_t, c_ptr, mistra_fastkmt_args
       integer(c_int64_t), value :: ncell
       type(mistra_fastkmt_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_fastkmt
  end interface
end module mistra_fastkmt_mod

! mistra_difc_mod - ISO_C_BINDING interface of include/mistra_difc.h: SUBROUTINE difc
! (str.f90:3271-3445) on s1, s3, sl1 and sion1 in place; ncol = 1 for the reference's single column.
module mistra_difc_mod
  use, intrinsic :: iso_c_binding
  implicit none
  integer, parameter :: MISTRA_DIFC_MAXFIELDS = 8
  type, bind(C) :: mistra_difc_field
     type(c_ptr) :: s
     integer(c_int32_t) :: row, nproc
  end type mistra_difc_field
  type, bind(C) :: mistra_difc_args
     integer(c_int32_t) :: n, nfield
     real(c_double) :: dt
     type(c_ptr) :: atkh, w, am3, detw, deta
     type(mistra_difc_field) :: field(MISTRA_DIFC_MAXFIELDS)
  end type mistra_difc_args
  interface
     function mistra_difc(ncol, a, stream) result(rc) bind(C, name="mistra_difc")
       import :: c_int, c_int64_t, c_ptr, mistra_difc_args
       integer(c_int64_t), value :: ncol
       type(mistra_difc_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_difc
  end interface
end module mistra_difc_mod

! mistra_drive_mod - ISO_C_BINDING interface of include/mistra_drive.h: the gather / scatter halves of
! gas_drive / aer_drive / tot_drive (aer.f:146-178, 233-245) for a batch of layers on device-resident arrays.
module mistra_drive_mod
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: mistra_drive_args
     integer(c_int32_t) :: nvar, nfix, j1, j5, j2, j6, nkc, nmap
     type(c_ptr) :: map_kpp, map_arr, map_off, layer, s1, s3, sl1, sion1, var, fix
     integer(c_int32_t) :: clamp_liquid, f32_literals, indf_o2, indf_h2o, indf_n2
     integer(c_int32_t) :: indf_h2ol(4)
     type(c_ptr) :: air, h2o, cvv
     integer(c_int32_t) :: clip_negative
  end type mistra_drive_args
  interface
     function mistra_drive_gather_device(ncell, a, stream) result(rc) bind(C, name="mistra_drive_gather_device")
       import :: c_int, c_int64_t, c_ptr, mistra_drive_args
       integer(c_int64_t), value :: ncell
       type(mistra_drive_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_drive_gather_device
     function mistra_drive_scatter_device(ncell, a, stream) result(rc) bind(C, name="mistra_drive_scatter_device")
       import :: c_int, c_int64_t, c_ptr, mistra_drive_args
       integer(c_int64_t), value :: ncell
       type(mistra_drive_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_drive_scatter_device
  end interface
end module mistra_drive_mod

! mistra_sed_mod - ISO_C_BINDING interface of include/mistra_sed.h: SUBROUTINE sedp (str.f90:2257-2411),
! sedl (2627-2787) and the species loop of sedc (2567-2596) in place; ncol = 1 for the reference's single column.
module mistra_sed_mod
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: mistra_sedp_args
     integer(c_int32_t) :: n, nf, nka, nkt
     real(c_double) :: dt
     type(c_ptr) :: detw, deta, t, p, rq, e, kw, vd, ff, diag
  end type mistra_sedp_args
  type, bind(C) :: mistra_sedl_args
     integer(c_int32_t) :: n, nf, nkc, nkc_l, j2, j6
     real(c_double) :: dt
     type(c_ptr) :: detw, deta, t, p, rc, vt, vdm, sl1, sion1
  end type mistra_sedl_args
  type, bind(C) :: mistra_sedc_args
     integer(c_int32_t) :: n, j1
     real(c_double) :: dt
     type(c_ptr) :: detw, deta, vg, es1, s1
  end type mistra_sedc_args
  interface
     function mistra_sedp(ncol, a, stream) result(rc) bind(C, name="mistra_sedp")
       import :: c_int, c_int64_t, c_ptr, mistra_sedp_args
       integer(c_int64_t), value :: ncol
       type(mistra_sedp_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_sedp
     function mistra_sedl(ncol, a, stream) result(rc) bind(C, name="mistra_sedl")
       import :: c_int, c_int64_t, c_ptr, mistra_sedl_args
       integer(c_int64_t), value :: ncol
       type(mistra_sedl_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_sedl
     function mistra_sedc(ncol, a, stream) result(rc) bind(C, name="mistra_sedc")
       import :: c_int, c_int64_t, c_ptr, mistra_sedc_args
       integer(c_int64_t), value :: ncol
       type(mistra_sedc_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_sedc
  end interface
end module mistra_sed_mod

! mistra_driver_mod - ISO_C_BINDING interface of include/mistra_driver.h: the layer loop of kpp_driver
! (kpp.f90:4305-4470) for an ensemble of columns: per-layer scalars, switches and the layers sorted by mechanism.
module mistra_driver_mod
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: mistra_driver_args
     integer(c_int32_t) :: n, nf, nkc, nphrxn
     integer(c_int32_t) :: halo, iod, lpBuys13_0D, neula
     integer(c_int32_t) :: box, n_bl
     integer(c_int32_t) :: kinv, nadv, j1, j5
     real(c_double) :: dt_ch
     type(c_ptr) :: u0, t, p, rho, cm3, am3, xm1, conv2, cm, cloud, photol_j, adv_row, xadv, s1, s3
     type(c_ptr) :: cb1, scal, ph_rat, air, h2o, cvv, mech, layers, count
  end type mistra_driver_args
  interface
     function mistra_driver_layers(ncol, a, stream) result(rc) bind(C, name="mistra_driver_layers")
       import :: c_int, c_int64_t, c_ptr, mistra_driver_args
       integer(c_int64_t), value :: ncol
       type(mistra_driver_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_driver_layers
     function mistra_driver_layers_device(ncol, a, stream) result(rc) bind(C, name="mistra_driver_layers_device")
       import :: c_int, c_int64_t, c_ptr, mistra_driver_args
       integer(c_int64_t), value :: ncol
       type(mistra_driver_args), intent(in) :: a
       type(c_ptr), value :: stream
       integer(c_int) :: rc
     end function mistra_driver_layers_device
  end interface
end module mistra_driver_mod
