!! libGPU -- iso_c_binding shim between VOLCANOR's Fortran driver and libvolcanor_b200.so
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (DESIGN.md, "Boundary").
!! The call sequence below is the one tests/case_hooks.py:gpu_hooks drives against the same C ABI from the
!! C restatement of `program main` (oracle/vlc_case.c), where it is tested on a B200 (tests/test_gpu_case.py).
!!
!! Build (reference tree, with a Fortran compiler):
!!   gfortran -c libGPU.f90   (after classdef.f90; add to CMakeLists.txt next to libCommon.f90)
!!   link: -L<repo>/volcanor_b200 -lvolcanor_b200 -lcudart -lcusolver
!!
!! What changes in the driver (INTEGRATION.md has the full diff):
!!   libCommon.f90: `use libGPU`, and the bodies of vind_onNwake_byRotor / vind_onFwake_byRotor become one call each
!!   main.f90     : the three OpenMP CP loops (RHS :528-573, initial :124-167, forces :632-656) call
!!                  gpu_vind_points on the array of collocation points instead of rotor%vind_* per point;
!!                  calcAIC / matmulAX(AIC_inv, RHS) become gpu_calcAIC / gpu_solve
!! Nothing else (case files, derived types, time integration, loads, output) is touched.
module libGPU
  use, intrinsic :: iso_c_binding
  use classdef, only: rotor_class, Nwake_class, Fwake_class, dp
  implicit none
  private
  public :: gpu_init, gpu_finalize, gpu_sync_rotor, gpu_vind_onNwake_byRotor, gpu_vind_onFwake_byRotor
  public :: gpu_vind_points, gpu_calcAIC, gpu_solve, gpu_touch
  ! device-resident time stepping (tier 2b of the C ABI): the wake is uploaded once and every wake mutator of the time
  ! loop runs on the library's copies; tests/native/case_gpu_hooks.c (resident mode) is the tested C twin
  public :: gpu_resident_begin, gpu_wake_prestep, gpu_wake_convect, gpu_download_wake
  public :: gpu_burst_wake, gpu_calc_skew
  ! collocation-point stage on the device (tier 2c of the C ABI): velCP, RHS, solve, map_gam and -- forceCalcSwitch 0 --
  ! velCPTotal and the sectional loads; tests/native/case_gpu_hooks.c (h_cp_rhs_solve / h_cp_forces) is the tested C twin
  public :: gpu_cp_rhs_solve, gpu_cp_forces
  logical, save :: resident = .false.
  integer(c_int), parameter :: VEL_FIRST_STEP = 0, VEL_AB2 = 1, VEL_AM2 = 2, VEL_SHIFT_HISTORY = 3, VEL_ORDER2 = 4, &
    & VEL_COPY_TO_STEP = 5
  ! velocity arrays by id (vlc_rotor_wakevel_copy / _lincomb)
  integer(c_int), parameter :: ARR_VEL = 0, ARR_VEL1 = 1, ARR_PREDICTED = 2, ARR_STEP = 3, ARR_VEL2 = 4, ARR_VEL3 = 5

  type(c_ptr), save :: ctx = c_null_ptr

  ! What the library currently holds is stale for rotor ir: the wing, the current ('C') wake, the predicted ('P') wake.
  ! Everything starts stale; gpu_sync_rotor clears the flags it serves; the driver calls gpu_touch after it changes
  ! state (main.f90: after move/rot_advance and map_gam -> GPU_WING; after assignshed/age_wake/dissipate_wake/
  ! convectwake('C')/strain_wake/rollup -> GPU_WAKE_C; after the copy to *Predicted and convectwake('P') -> GPU_WAKE_P).
  ! Without any gpu_touch calls every sweep re-sends everything (correct, slower): tests/native/case_gpu_hooks.c
  ! measures 0.99 s -> 0.21 s for the 160 steps of katzNplotkin-AR04 with the flags.
  integer, parameter, public :: GPU_WING = 1, GPU_WAKE_C = 2, GPU_WAKE_P = 3
  logical, allocatable, save :: stale(:, :)
  logical, save :: trackStale = .false.

  ! what for gpu_vind_points
  integer, parameter, public :: GPU_BYWING = 0, GPU_BYWAKE = 1, GPU_BOTH = 2, GPU_BOUNDVORTICES = 3

  interface
    integer(c_int) function vlc_create(device, out) bind(C, name='vlc_create')
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: out
    end function
    integer(c_int) function vlc_create_multi(n_devices, devices, out) bind(C, name='vlc_create_multi')
      import :: c_int, c_ptr
      integer(c_int), value :: n_devices
      integer(c_int), intent(in) :: devices(*)
      type(c_ptr), intent(out) :: out
    end function
    integer(c_int) function vlc_rotors_clear(c) bind(C, name='vlc_rotors_clear')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
    end function
    integer(c_int) function vlc_destroy(c) bind(C, name='vlc_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
    end function
    type(c_ptr) function vlc_last_error(c) bind(C, name='vlc_last_error')
      import :: c_ptr
      type(c_ptr), value :: c
    end function
    integer(c_int) function vlc_rotor_define(c, ir, nb, nc, ns, nNwake, nFwake, surfaceType) &
        & bind(C, name='vlc_rotor_define')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir, nb, nc, ns, nNwake, nFwake, surfaceType
    end function
    integer(c_int) function vlc_rotor_set_rows(c, ir, rowNear, rowFar) bind(C, name='vlc_rotor_set_rows')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir, rowNear, rowFar
    end function
    integer(c_int) function vlc_rotor_put_wing(c, ir, ib, wiP) bind(C, name='vlc_rotor_put_wing')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib
      real(c_double), intent(in) :: wiP(*)
    end function
    integer(c_int) function vlc_rotor_put_nwake(c, ir, ib, predicted, waN) bind(C, name='vlc_rotor_put_nwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(in) :: waN(*)
    end function
    integer(c_int) function vlc_rotor_put_fwake(c, ir, ib, predicted, waF) bind(C, name='vlc_rotor_put_fwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(in) :: waF(*)
    end function
    integer(c_int) function vlc_rotor_put_pfwake(c, ir, ib, predicted, wapF) bind(C, name='vlc_rotor_put_pfwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(in) :: wapF(*)
    end function
    integer(c_int) function vlc_rotor_vind_bywing(c, ir, m, P, V) bind(C, name='vlc_rotor_vind_bywing')
      import :: c_int, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      integer(c_int64_t), value :: m
      real(c_double), intent(in) :: P(3, *)
      real(c_double), intent(out) :: V(3, *)
    end function
    integer(c_int) function vlc_rotor_vind_bywake(c, ir, predicted, m, P, V) bind(C, name='vlc_rotor_vind_bywake')
      import :: c_int, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, predicted
      integer(c_int64_t), value :: m
      real(c_double), intent(in) :: P(3, *)
      real(c_double), intent(out) :: V(3, *)
    end function
    integer(c_int) function vlc_rotor_vind(c, ir, predicted, m, P, V) bind(C, name='vlc_rotor_vind')
      import :: c_int, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, predicted
      integer(c_int64_t), value :: m
      real(c_double), intent(in) :: P(3, *)
      real(c_double), intent(out) :: V(3, *)
    end function
    integer(c_int) function vlc_rotor_vind_bywing_boundVortices(c, ir, m, P, V) &
        & bind(C, name='vlc_rotor_vind_bywing_boundVortices')
      import :: c_int, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      integer(c_int64_t), value :: m
      real(c_double), intent(in) :: P(3, *)
      real(c_double), intent(out) :: V(3, *)
    end function
    integer(c_int) function vlc_vind_onNwake_byRotor(c, ir, Nwake, rows, cols, ld, predicted, vindArray) &
        & bind(C, name='vlc_vind_onNwake_byRotor')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, rows, cols, ld, predicted
      real(c_double), intent(in) :: Nwake(*)
      real(c_double), intent(out) :: vindArray(3, rows, cols + 1)
    end function
    integer(c_int) function vlc_vind_onFwake_byRotor(c, ir, Fwake, rows, predicted, vindArray) &
        & bind(C, name='vlc_vind_onFwake_byRotor')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, rows, predicted
      real(c_double), intent(in) :: Fwake(*)
      real(c_double), intent(out) :: vindArray(3, rows)
    end function
    integer(c_int) function vlc_rotor_calcAIC(c, ir, AIC_out) bind(C, name='vlc_rotor_calcAIC')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), intent(out) :: AIC_out(*)
    end function
    integer(c_int) function vlc_rotor_solve(c, ir, RHS, gamVec) bind(C, name='vlc_rotor_solve')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), intent(in) :: RHS(*)
      real(c_double), intent(out) :: gamVec(*)
    end function
    ! ---- tier 2b: the reference's wake mutators on the device copies ----
    integer(c_int) function vlc_rotor_set_wake_params(c, ir, nbConvect, axisymmetrySwitch, ductSwitch, &
        & suppressFwakeSwitch, rollupStart, rollupEnd, rollupSign, apparentViscCoeff, decayCoeff, initWakeVel) &
        & bind(C, name='vlc_rotor_set_wake_params')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, nbConvect, axisymmetrySwitch, ductSwitch, suppressFwakeSwitch, rollupStart, rollupEnd
      real(c_double), value :: rollupSign, apparentViscCoeff, decayCoeff, initWakeVel
    end function
    integer(c_int) function vlc_rotor_set_frame(c, ir, shaftAxis, hubCoords) bind(C, name='vlc_rotor_set_frame')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), intent(in) :: shaftAxis(3), hubCoords(3)
    end function
    integer(c_int) function vlc_rotor_assignshed(c, ir, edge) bind(C, name='vlc_rotor_assignshed')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir, edge
    end function
    integer(c_int) function vlc_rotor_age_wake(c, ir, dt, omegaSlow) bind(C, name='vlc_rotor_age_wake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), value :: dt, omegaSlow
    end function
    integer(c_int) function vlc_rotor_dissipate_wake(c, ir, dt, kinematicVisc) bind(C, name='vlc_rotor_dissipate_wake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), value :: dt, kinematicVisc
    end function
    integer(c_int) function vlc_rotor_strain_wake(c, ir) bind(C, name='vlc_rotor_strain_wake')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_rotor_wake_to_predicted(c, ir) bind(C, name='vlc_rotor_wake_to_predicted')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_rotor_convectwake(c, ir, dt, predicted) bind(C, name='vlc_rotor_convectwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, predicted
      real(c_double), value :: dt
    end function
    integer(c_int) function vlc_rotor_rollup(c, ir) bind(C, name='vlc_rotor_rollup')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_wake_sweep(c, predicted, addInitWakeVel) bind(C, name='vlc_wake_sweep')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: predicted, addInitWakeVel
    end function
    integer(c_int) function vlc_rotor_wakevel_op(c, ir, op) bind(C, name='vlc_rotor_wakevel_op')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir, op
    end function
    integer(c_int) function vlc_rotor_get_nwake(c, ir, ib, predicted, waN) bind(C, name='vlc_rotor_get_nwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(out) :: waN(*)
    end function
    integer(c_int) function vlc_rotor_get_fwake(c, ir, ib, predicted, waF) bind(C, name='vlc_rotor_get_fwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(out) :: waF(*)
    end function
    integer(c_int) function vlc_rotor_put_pfwake_helix(c, ir, ib, predicted, helix) bind(C, name='vlc_rotor_put_pfwake_helix')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(in) :: helix(*)
    end function
    integer(c_int) function vlc_rotor_put_wakevel(c, ir, ib, which, velN, velF) bind(C, name='vlc_rotor_put_wakevel')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, which
      real(c_double), intent(in) :: velN(*)
      real(c_double), intent(in) :: velF(*)
    end function
    integer(c_int) function vlc_rotor_get_wakevel(c, ir, ib, which, velN, velF) bind(C, name='vlc_rotor_get_wakevel')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, which
      real(c_double), intent(out) :: velN(*)
      real(c_double), intent(out) :: velF(*)
    end function
    integer(c_int) function vlc_rotor_calc_skew(c, ir) bind(C, name='vlc_rotor_calc_skew')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_rotor_burst_wake(c, ir, skewLimit, largeCoreRadius) bind(C, name='vlc_rotor_burst_wake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), value :: skewLimit, largeCoreRadius
    end function
    integer(c_int) function vlc_rotor_updatePrescribedWake(c, ir, deltaPsi, prescWakeGenNt, predicted) &
      & bind(C, name='vlc_rotor_updatePrescribedWake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), value :: deltaPsi
      integer(c_int), value :: prescWakeGenNt, predicted
    end function
    integer(c_int) function vlc_rotor_get_pfwake(c, ir, ib, predicted, wapF, helix) bind(C, name='vlc_rotor_get_pfwake')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib, predicted
      real(c_double), intent(out) :: wapF(*)
      real(c_double), intent(out) :: helix(*)
    end function
    integer(c_int) function vlc_rotor_wakevel_copy(c, ir, dst, src) bind(C, name='vlc_rotor_wakevel_copy')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir, dst, src
    end function
    integer(c_int) function vlc_rotor_wakevel_lincomb(c, ir, dst, nterms, src, coef, divisor) &
        & bind(C, name='vlc_rotor_wakevel_lincomb')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, dst, nterms
      integer(c_int), intent(in) :: src(*)
      real(c_double), intent(in) :: coef(*)
      real(c_double), value :: divisor
    end function
    ! tier 2c
    integer(c_int) function vlc_rotor_reset_velCP(c, ir) bind(C, name='vlc_rotor_reset_velCP')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_rotor_calc_RHS(c, ir, velCP_out, RHS_out) bind(C, name='vlc_rotor_calc_RHS')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), intent(out) :: velCP_out(3, *), RHS_out(*)
    end function
    integer(c_int) function vlc_rotor_solve_map_gam(c, ir, gamVec_out) bind(C, name='vlc_rotor_solve_map_gam')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir
      real(c_double), intent(out) :: gamVec_out(*)
    end function
    integer(c_int) function vlc_rotor_put_sections(c, ir, ib, sec) bind(C, name='vlc_rotor_put_sections')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib
      real(c_double), intent(in) :: sec(*)
    end function
    integer(c_int) function vlc_rotor_calc_velCPTotal(c, ir) bind(C, name='vlc_rotor_calc_velCPTotal')
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int), value :: ir
    end function
    integer(c_int) function vlc_rotor_calc_force(c, ir, density, dt, Omega, spanwiseLiftSwitch) &
        & bind(C, name='vlc_rotor_calc_force')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, spanwiseLiftSwitch
      real(c_double), value :: density, dt, Omega
    end function
    integer(c_int) function vlc_rotor_get_loads(c, ir, ib, loads) bind(C, name='vlc_rotor_get_loads')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib
      real(c_double), intent(out) :: loads(*)
    end function
    integer(c_int) function vlc_rotor_get_wing(c, ir, ib, wiP) bind(C, name='vlc_rotor_get_wing')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: c
      integer(c_int), value :: ir, ib
      real(c_double), intent(out) :: wiP(*)
    end function
  end interface

contains

  subroutine check(rc)
    !! The library returns a status; the reference aborts with `error stop` (e.g. libMath.f90:73).
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc /= 0) then
      call c_f_pointer(vlc_last_error(ctx), msg, [512])
      n = 1
      do while (n < 512 .and. msg(n) /= c_null_char)
        n = n + 1
      enddo
      print *, 'volcanor_b200: ', msg(1:n - 1)
      error stop 'ERROR: GPU library call failed'
    endif
  end subroutine check

  subroutine gpu_init(rotor, device)
    !! Once, after rotor%init (main.f90:31-40): declare every rotor's sizes.
    !! VOLCANOR_GPUS=n in the environment (n > 1) makes the ONE handle span the GPUs 0 .. n-1 of the box
    !! (vlc_create_multi): the library replicates the wake, shards the targets of every sweep and all-gathers the
    !! velocity slices with NCCL; no call site below changes (main.f90:814-841 keeps calling vind_onNwake_byRotor).
    type(rotor_class), intent(in) :: rotor(:)
    integer, intent(in) :: device
    integer :: ir, ngpu, stat
    integer(c_int), allocatable :: devlist(:)
    character(len=16) :: env
    ngpu = 1
    call get_environment_variable('VOLCANOR_GPUS', env, status=stat)
    if (stat == 0) read (env, *, iostat=stat) ngpu
    if (stat /= 0) ngpu = 1
    if (ngpu > 1) then
      allocate (devlist(ngpu))
      devlist = [(int(ir - 1, c_int), ir=1, ngpu)]
      call check(vlc_create_multi(int(ngpu, c_int), devlist, ctx))
    else
      call check(vlc_create(int(device, c_int), ctx))
    endif
    call check(vlc_rotors_clear(ctx))
    allocate (stale(3, size(rotor)))
    stale = .true.
    do ir = 1, size(rotor)
      call check(vlc_rotor_define(ctx, ir - 1, rotor(ir)%nb, rotor(ir)%nc, rotor(ir)%ns, &
        & rotor(ir)%nNwake, rotor(ir)%nFwake, rotor(ir)%surfaceType))
    enddo
  end subroutine gpu_init

  subroutine gpu_finalize()
    if (c_associated(ctx)) call check(vlc_destroy(ctx))
    ctx = c_null_ptr
  end subroutine gpu_finalize

  subroutine gpu_touch(ir, what)
    !! The driver changed rotor ir's wing / 'C' wake / 'P' wake: send it again before the next sweep that reads it.
    integer, intent(in) :: ir, what
    trackStale = .true.
    stale(what, ir) = .true.
  end subroutine gpu_touch

  subroutine gpu_sync_rotor(rotor, ir, predicted)
    !! Flatten the non-interoperable derived types into arrays of doubles and upload them.
    !! vr_class = 50, Fwake_class = 13, wingpanel_class = 104 doubles (sequence of default reals(dp));
    !! `transfer` keeps the component order of classdef.f90:57-220.
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir
    logical, intent(in) :: predicted
    integer :: ib
    integer(c_int) :: p
    real(c_double), allocatable :: buf(:)
    logical :: sendWing, sendWake
    p = merge(1_c_int, 0_c_int, predicted)
    call check(vlc_rotor_set_rows(ctx, ir - 1, rotor%rowNear, rotor%rowFar))
    sendWing = stale(GPU_WING, ir) .or. .not. trackStale
    sendWake = stale(merge(GPU_WAKE_P, GPU_WAKE_C, predicted), ir) .or. .not. trackStale
    if (resident) then   ! the device's wake is the wake: only the frame and the wing travel
      sendWake = .false.
      call check(vlc_rotor_set_frame(ctx, ir - 1, rotor%shaftAxis, rotor%hubCoords))
    endif
    do ib = 1, rotor%nb
      if (sendWing) then
        buf = transfer(rotor%blade(ib)%wiP, buf)
        call check(vlc_rotor_put_wing(ctx, ir - 1, ib - 1, buf))
      endif
      if (rotor%nNwake > 0 .and. sendWake) then
        if (predicted) then
          buf = transfer(rotor%blade(ib)%waNPredicted, buf)
        else
          buf = transfer(rotor%blade(ib)%waN, buf)
        endif
        call check(vlc_rotor_put_nwake(ctx, ir - 1, ib - 1, p, buf))
        if (rotor%nFwake > 0) then
          if (predicted) then
            buf = transfer(rotor%blade(ib)%waFPredicted, buf)
          else
            buf = transfer(rotor%blade(ib)%waF, buf)
          endif
          call check(vlc_rotor_put_fwake(ctx, ir - 1, ib - 1, p, buf))
        endif
        if (rotor%prescWakeNt > 0) then
          if (predicted) then
            buf = transfer(rotor%blade(ib)%wapFPredicted%waF, buf)
          else
            buf = transfer(rotor%blade(ib)%wapF%waF, buf)
          endif
          call check(vlc_rotor_put_pfwake(ctx, ir - 1, ib - 1, p, buf))
        endif
      endif
    enddo
    if (sendWing) stale(GPU_WING, ir) = .false.
    if (sendWake) stale(merge(GPU_WAKE_P, GPU_WAKE_C, predicted), ir) = .false.
  end subroutine gpu_sync_rotor

  function gpu_vind_onNwake_byRotor(rotor, ir, Nwake, optionalChar) result(vindArray)
    !! Drop-in body of libCommon.f90:114-171 (ir = index of `rotor` in the global rotor array).
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir
    type(Nwake_class), intent(in), dimension(:, :) :: Nwake
    character(len=1), optional :: optionalChar
    real(dp), dimension(3, size(Nwake, 1), size(Nwake, 2) + 1) :: vindArray
    real(c_double), allocatable :: buf(:)
    logical :: pred
    pred = .false.
    if (present(optionalChar)) then
      if (optionalChar /= 'P') error stop 'ERROR: Wrong character flag for vind_onNwake_byRotor()'
      pred = .true.
    endif
    call gpu_sync_rotor(rotor, ir, pred)
    buf = transfer(Nwake, buf)           ! contiguous copy of the slice: ld = rows
    call check(vlc_vind_onNwake_byRotor(ctx, ir - 1, buf, size(Nwake, 1), size(Nwake, 2), size(Nwake, 1), &
      & merge(1_c_int, 0_c_int, pred), vindArray))
  end function gpu_vind_onNwake_byRotor

  function gpu_vind_onFwake_byRotor(rotor, ir, Fwake, optionalChar) result(vindArray)
    !! Drop-in body of libCommon.f90:173-211
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir
    type(Fwake_class), intent(in), dimension(:) :: Fwake
    character(len=1), optional :: optionalChar
    real(dp), dimension(3, size(Fwake)) :: vindArray
    real(c_double), allocatable :: buf(:)
    logical :: pred
    pred = .false.
    if (present(optionalChar)) then
      if (optionalChar /= 'P') error stop 'ERROR: Wrong character flag for vind_onFwake_byRotor()'
      pred = .true.
    endif
    if (size(Fwake) == 0) return
    call gpu_sync_rotor(rotor, ir, pred)
    buf = transfer(Fwake, buf)
    call check(vlc_vind_onFwake_byRotor(ctx, ir - 1, buf, size(Fwake), merge(1_c_int, 0_c_int, pred), vindArray))
  end function gpu_vind_onFwake_byRotor

  function gpu_vind_points(rotor, ir, what, P, predicted) result(V)
    !! rotor%vind_bywing / %vind_bywake / both / %vind_bywing_boundVortices at all points of P(3, m) in one call:
    !! replaces the per-point calls inside the OpenMP loops of main.f90:124-167, :247-271, :528-573, :632-656.
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir, what
    real(dp), intent(in) :: P(:, :)
    logical, intent(in) :: predicted
    real(dp) :: V(3, size(P, 2))
    integer(c_int64_t) :: m
    integer(c_int) :: pr
    m = size(P, 2, kind=c_int64_t)
    pr = merge(1_c_int, 0_c_int, predicted)
    call gpu_sync_rotor(rotor, ir, predicted)
    select case (what)
    case (GPU_BYWING)
      call check(vlc_rotor_vind_bywing(ctx, ir - 1, m, P, V))
    case (GPU_BYWAKE)
      call check(vlc_rotor_vind_bywake(ctx, ir - 1, pr, m, P, V))
    case (GPU_BOTH)
      call check(vlc_rotor_vind(ctx, ir - 1, pr, m, P, V))
    case (GPU_BOUNDVORTICES)
      call check(vlc_rotor_vind_bywing_boundVortices(ctx, ir - 1, m, P, V))
    case default
      error stop 'ERROR: gpu_vind_points: wrong selector'
    end select
  end function gpu_vind_points

  subroutine gpu_calcAIC(rotor, ir)
    !! rotor%calcAIC() (classdef.f90:4151-4179): assemble on the device, LU-factor once (cuSOLVER getrf).
    !! rotor%AIC is filled for output / isInverse-style checks; AIC_inv is not needed any more.
    type(rotor_class), intent(inout) :: rotor
    integer, intent(in) :: ir
    call gpu_sync_rotor(rotor, ir, .false.)
    call check(vlc_rotor_calcAIC(ctx, ir - 1, rotor%AIC))
  end subroutine gpu_calcAIC

  function gpu_solve(rotor, ir) result(gamVec)
    !! gamVec = matmulAX(AIC_inv, RHS) (main.f90:190, :596) -> getrs with the stored factors
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir
    real(dp) :: gamVec(size(rotor%RHS))
    call check(vlc_rotor_solve(ctx, ir - 1, rotor%RHS, gamVec))
  end function gpu_solve

  ! ------------------------------------------------------------------ device-resident time stepping (tier 2b)

  subroutine gpu_resident_begin(rotor)
    !! Once, before the time loop (after main.f90:227-234): the wake records as rotor%init and the first
    !! assignshed('TE') left them go up whole, current and predicted; from here on the driver must not touch
    !! waN / waF / waNPredicted / waFPredicted / velNwake* / velFwake* (gpu_download_wake brings them back for plots).
    type(rotor_class), intent(in) :: rotor(:)
    integer :: ir, ib
    real(c_double), allocatable :: buf(:)
    do ir = 1, size(rotor)
      call check(vlc_rotor_set_wake_params(ctx, ir - 1, rotor(ir)%nbConvect, rotor(ir)%axisymmetrySwitch, &
        & rotor(ir)%ductSwitch, rotor(ir)%suppressFwakeSwitch, rotor(ir)%rollupStart, rotor(ir)%rollupEnd, &
        & rotor(ir)%Omega*rotor(ir)%controlPitch(1), rotor(ir)%apparentViscCoeff, rotor(ir)%decayCoeff, &
        & rotor(ir)%initWakeVel))
      call check(vlc_rotor_set_rows(ctx, ir - 1, 1, 1))   ! every row travels this once
      do ib = 1, rotor(ir)%nb
        if (rotor(ir)%nNwake > 0) then
          buf = transfer(rotor(ir)%blade(ib)%waN, buf)
          call check(vlc_rotor_put_nwake(ctx, ir - 1, ib - 1, 0_c_int, buf))
          if (allocated(rotor(ir)%blade(ib)%waNPredicted)) then   ! fdScheme 1, 3, 4, 5 (classdef.f90:3733-3824)
            buf = transfer(rotor(ir)%blade(ib)%waNPredicted, buf)
            call check(vlc_rotor_put_nwake(ctx, ir - 1, ib - 1, 1_c_int, buf))
          endif
          ! zero at a fresh start like the device's own; the histories of a run resumed from a restart file
          call put_vel(ir, ib, ARR_VEL, rotor(ir)%blade(ib)%velNwake, rotor(ir)%blade(ib)%velFwake)
          call put_vel(ir, ib, ARR_VEL1, rotor(ir)%blade(ib)%velNwake1, rotor(ir)%blade(ib)%velFwake1)
          call put_vel(ir, ib, ARR_PREDICTED, rotor(ir)%blade(ib)%velNwakePredicted, rotor(ir)%blade(ib)%velFwakePredicted)
          call put_vel(ir, ib, ARR_STEP, rotor(ir)%blade(ib)%velNwakeStep, rotor(ir)%blade(ib)%velFwakeStep)
          call put_vel(ir, ib, ARR_VEL2, rotor(ir)%blade(ib)%velNwake2, rotor(ir)%blade(ib)%velFwake2)
          call put_vel(ir, ib, ARR_VEL3, rotor(ir)%blade(ib)%velNwake3, rotor(ir)%blade(ib)%velFwake3)
        endif
        if (rotor(ir)%nFwake > 0) then
          buf = transfer(rotor(ir)%blade(ib)%waF, buf)
          call check(vlc_rotor_put_fwake(ctx, ir - 1, ib - 1, 0_c_int, buf))
          if (allocated(rotor(ir)%blade(ib)%waFPredicted)) then
            buf = transfer(rotor(ir)%blade(ib)%waFPredicted, buf)
            call check(vlc_rotor_put_fwake(ctx, ir - 1, ib - 1, 1_c_int, buf))
          endif
        endif
        if (rotor(ir)%prescWakeNt > 0) then   ! all zero at a fresh start; the helix and its fit of a resumed run
          buf = transfer(rotor(ir)%blade(ib)%wapF%waF, buf)
          call check(vlc_rotor_put_pfwake(ctx, ir - 1, ib - 1, 0_c_int, buf))
          call check(vlc_rotor_put_pfwake_helix(ctx, ir - 1, ib - 1, 0_c_int, &
            & [rotor(ir)%blade(ib)%wapF%helixPitch, rotor(ir)%blade(ib)%wapF%helixRadius]))
          buf = transfer(rotor(ir)%blade(ib)%wapFPredicted%waF, buf)
          call check(vlc_rotor_put_pfwake(ctx, ir - 1, ib - 1, 1_c_int, buf))
          call check(vlc_rotor_put_pfwake_helix(ctx, ir - 1, ib - 1, 1_c_int, &
            & [rotor(ir)%blade(ib)%wapFPredicted%helixPitch, rotor(ir)%blade(ib)%wapFPredicted%helixRadius]))
        endif
      enddo
    enddo
    resident = .true.
    trackStale = .true.
  end subroutine gpu_resident_begin

  subroutine gpu_wake_prestep(rotor, dt, wakeDissipation, kinematicVisc)
    !! Replaces main.f90:466-506: assignshed('LE'), age_wake, dissipate_wake of every rotor, in the driver's order.
    !! Call gpu_touch(ir, GPU_WING) after the driver moves a rotor (main.f90:455-463) and after map_gam.
    type(rotor_class), intent(in) :: rotor(:)
    real(dp), intent(in) :: dt, kinematicVisc
    integer, intent(in) :: wakeDissipation
    integer :: ir
    do ir = 1, size(rotor)
      call gpu_sync_rotor(rotor(ir), ir, .false.)
    enddo
    do ir = 1, size(rotor)
      call check(vlc_rotor_assignshed(ctx, ir - 1, 0_c_int))
    enddo
    do ir = 1, size(rotor)
      call check(vlc_rotor_age_wake(ctx, ir - 1, dt, rotor(ir)%omegaSlow))
    enddo
    if (wakeDissipation == 1) then
      do ir = 1, size(rotor)
        call check(vlc_rotor_dissipate_wake(ctx, ir - 1, dt, kinematicVisc))
      enddo
    endif
  end subroutine gpu_wake_prestep

  subroutine gpu_burst_wake(rotor, ir)
    !! Replaces `call rotor(ir)%burst_wake()` (main.f90:490-497; classdef.f90:4911-4917) in resident mode: the driver keeps
    !! its `mod(iter, switches%wakeBurst)` test and calls this right after gpu_wake_prestep.
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir
    call check(vlc_rotor_burst_wake(ctx, ir - 1, rotor%skewLimit, rotor%chord))
  end subroutine gpu_burst_wake

  subroutine gpu_calc_skew(ir)
    !! Replaces `call rotor(ir)%calc_skew()` (main.f90:499-504; classdef.f90:4919-4936) in resident mode; skew2file then
    !! needs the records on the host: gpu_download_wake.
    integer, intent(in) :: ir
    call check(vlc_rotor_calc_skew(ctx, ir - 1))
  end subroutine gpu_calc_skew

  subroutine gpu_convect(rotor, ir, iter, dt, p)
    !! rotor%convectwake(iter, dt, wakeType) (classdef.f90:4786-4830) on the device records, its last statement included:
    !! the prescribed far wake (:4826-4828, rotor%updatePrescribedWake :5170-5218) is regenerated on the device from the
    !! device's own far rows; gpu_download_wake brings the helix back for wake plots.
    type(rotor_class), intent(in) :: rotor
    integer, intent(in) :: ir, iter
    real(dp), intent(in) :: dt
    integer(c_int), intent(in) :: p
    call check(vlc_rotor_convectwake(ctx, ir - 1, dt, p))
    if (rotor%prescWakeNt > 0 .and. iter > rotor%prescWakeNt) then
      call check(vlc_rotor_updatePrescribedWake(ctx, ir - 1, rotor%omegaSlow*dt, int(rotor%prescWakeGenNt, c_int), p))
    endif
  end subroutine gpu_convect

  ! One MPI rank per GPU: replace each `vlc_wake_sweep(ctx, p, addInit)` below by
  !   vlc_wake_sweep_count(ctx, M); vlc_wake_sweep_slice(ctx, p, first, count, d_vel) on this rank's slice of the M targets;
  !   MPI_Allgather / ncclAllGather of d_vel (3*M doubles on the device, slices of ceiling(M/nranks) targets);
  !   vlc_wake_sweep_scatter(ctx, p, addInit, d_vel)
  ! -- every rank keeps the whole wake and runs the other stages redundantly (tests/multi_gpu_case.py is the tested twin).
  subroutine gpu_wake_convect(rotor, iter, dt, fdScheme, wakeStrain, initWakeVelNt)
    !! Replaces main.f90:800-1440: the wake sweeps, the fdScheme switch (0 explicit Euler :846-859, 1 predictor-
    !! corrector :861-949, 2 explicit Adams-Bashforth :951-1000, 3 Adams-Bashforth / Adams-Moulton :1002-1115, 4 and 5 the
    !! same of third and fourth order :1117-1404) with its velocity bookkeeping, strain_wake, rollup, assignshed('TE').
    type(rotor_class), intent(in) :: rotor(:)
    integer, intent(in) :: iter, fdScheme, wakeStrain, initWakeVelNt
    real(dp), intent(in) :: dt
    integer :: ir, start
    integer(c_int) :: addInit
    integer(c_int), parameter :: hist(3) = [ARR_VEL1, ARR_VEL2, ARR_VEL3]
    addInit = merge(1_c_int, 0_c_int, iter < initWakeVelNt)
    do ir = 1, size(rotor)
      call gpu_sync_rotor(rotor(ir), ir, .false.)   ! the solve changed the wing's circulation
    enddo
    call check(vlc_wake_sweep(ctx, 0_c_int, addInit))
    select case (fdScheme)
    case (0)
      do ir = 1, size(rotor)
        call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
      enddo
    case (2)   ! explicit Adams-Bashforth (:951-1000): velStep = vel1 = vel = 0.5*(3*vel - vel1), then convect
      do ir = 1, size(rotor)
        if (iter == 1) then
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_FIRST_STEP))
        else
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_AB2))
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_FIRST_STEP))
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_COPY_TO_STEP))
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
        endif
      enddo
    case (1)
      do ir = 1, size(rotor)
        call check(vlc_rotor_wake_to_predicted(ctx, ir - 1))
        call gpu_convect(rotor(ir), ir, iter, dt, 1_c_int)
      enddo
      call check(vlc_wake_sweep(ctx, 1_c_int, addInit))
      do ir = 1, size(rotor)
        call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_ORDER2))
        call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
      enddo
    case (3)
      if (iter == 1) then
        do ir = 1, size(rotor)
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_FIRST_STEP))
        enddo
      else
        do ir = 1, size(rotor)
          call check(vlc_rotor_wake_to_predicted(ctx, ir - 1))
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_AB2))
          call gpu_convect(rotor(ir), ir, iter, dt, 1_c_int)
        enddo
        call check(vlc_wake_sweep(ctx, 1_c_int, addInit))
        do ir = 1, size(rotor)
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_AM2))
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
          call check(vlc_rotor_wakevel_op(ctx, ir - 1, VEL_SHIFT_HISTORY))
        enddo
      endif
    case (4, 5)   ! Adams-Bashforth / Adams-Moulton of third (:1117-1248) and fourth order (:1250-1404)
      if (fdScheme == 4) then        ! its `iter == 0` start branch never runs inside the time loop
        start = merge(2, 0, iter == 2)
      else
        start = merge(iter, 0, iter <= 3)
      endif
      if (start > 0) then            ! this step only fills the history vel1 / vel2 / vel3
        do ir = 1, size(rotor)
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
          call check(vlc_rotor_wakevel_copy(ctx, ir - 1, hist(start), ARR_VEL))
        enddo
      else
        do ir = 1, size(rotor)
          call check(vlc_rotor_wake_to_predicted(ctx, ir - 1))
          call check(vlc_rotor_wakevel_copy(ctx, ir - 1, ARR_STEP, ARR_VEL))
          if (fdScheme == 4) then
            call check(vlc_rotor_wakevel_lincomb(ctx, ir - 1, ARR_VEL, 3_c_int, [ARR_VEL, ARR_VEL2, ARR_VEL1], &
              & [23._c_double, -16._c_double, 5._c_double], 12._c_double))
          else
            call check(vlc_rotor_wakevel_lincomb(ctx, ir - 1, ARR_VEL, 4_c_int, [ARR_VEL, ARR_VEL3, ARR_VEL2, ARR_VEL1], &
              & [55._c_double, -59._c_double, 37._c_double, -9._c_double], 24._c_double))
          endif
          call gpu_convect(rotor(ir), ir, iter, dt, 1_c_int)
        enddo
        call check(vlc_wake_sweep(ctx, 1_c_int, addInit))
        do ir = 1, size(rotor)
          if (fdScheme == 4) then
            call check(vlc_rotor_wakevel_lincomb(ctx, ir - 1, ARR_VEL, 3_c_int, [ARR_PREDICTED, ARR_STEP, ARR_VEL2], &
              & [5._c_double, 8._c_double, -1._c_double], 12._c_double))
          else
            call check(vlc_rotor_wakevel_lincomb(ctx, ir - 1, ARR_VEL, 4_c_int, [ARR_PREDICTED, ARR_STEP, ARR_VEL3, ARR_VEL2], &
              & [9._c_double, 19._c_double, -5._c_double, 1._c_double], 24._c_double))
          endif
          call gpu_convect(rotor(ir), ir, iter, dt, 0_c_int)
          call check(vlc_rotor_wakevel_copy(ctx, ir - 1, ARR_VEL1, ARR_VEL2))
          if (fdScheme == 4) then
            call check(vlc_rotor_wakevel_copy(ctx, ir - 1, ARR_VEL2, ARR_STEP))
          else
            call check(vlc_rotor_wakevel_copy(ctx, ir - 1, ARR_VEL2, ARR_VEL3))
            call check(vlc_rotor_wakevel_copy(ctx, ir - 1, ARR_VEL3, ARR_STEP))
          endif
        enddo
      endif
    case default
      error stop 'ERROR: gpu_wake_convect: unknown fdScheme'
    end select
    if (wakeStrain == 1) then
      do ir = 1, size(rotor)
        call check(vlc_rotor_strain_wake(ctx, ir - 1))
      enddo
    endif
    do ir = 1, size(rotor)
      if (rotor(ir)%nNwake <= 0) cycle
      if (rotor(ir)%rowNear == 1) call check(vlc_rotor_rollup(ctx, ir - 1))
      call check(vlc_rotor_assignshed(ctx, ir - 1, 1_c_int))
    enddo
  end subroutine gpu_wake_convect

  ! ------------------------------------------------------------------ collocation-point stage (tier 2c)

  subroutine gpu_cp_rhs_solve(rotor, subIter)
    !! Replaces main.f90:548-603 for every rotor inside ONE pass of ntSubLoop (:522).  The driver keeps :528-547 -- it writes
    !! the kinematic part into wiP%velCP / velCPm -- saves gamVecPrev, calls gpu_touch(ir, GPU_WING) and then this routine;
    !! velCP, RHS, gamVec come back, the wing circulation is mapped on both sides (rotor%map_gam() here, map_gam on the
    !! device), so the wing is NOT stale afterwards.  subIter (optional) = the loop index i of ntSubLoop: for i >= 1 the
    !! wing did not move, so nothing is uploaded and the device restarts velCP from velCPm (vlc_rotor_reset_velCP) before
    !! it adds the induced velocities again -- with the other rotors' circulations of pass i-1, as in the reference.
    type(rotor_class), intent(inout) :: rotor(:)
    integer, intent(in), optional :: subIter
    integer :: ir, ib, ic, is, q, pass
    real(c_double), allocatable :: velCP(:, :)
    pass = 0
    if (present(subIter)) pass = subIter
    do ir = 1, size(rotor)
      if (pass == 0) then
        call gpu_sync_rotor(rotor(ir), ir, .false.)
      else
        call check(vlc_rotor_reset_velCP(ctx, ir - 1))
      endif
    enddo
    do ir = 1, size(rotor)
      allocate (velCP(3, rotor(ir)%nbConvect*rotor(ir)%nc*rotor(ir)%ns))
      call check(vlc_rotor_calc_RHS(ctx, ir - 1, velCP, rotor(ir)%RHS))
      q = 0
      do ib = 1, rotor(ir)%nbConvect      ! wiP order: ic fastest, then is, then the blade
        do is = 1, rotor(ir)%ns
          do ic = 1, rotor(ir)%nc
            q = q + 1
            rotor(ir)%blade(ib)%wiP(ic, is)%velCP = velCP(:, q)
          enddo
        enddo
      enddo
      deallocate (velCP)
    enddo
    do ir = 1, size(rotor)
      call check(vlc_rotor_solve_map_gam(ctx, ir - 1, rotor(ir)%gamVec))
      call rotor(ir)%map_gam()
    enddo
  end subroutine gpu_cp_rhs_solve

  subroutine gpu_cp_forces(rotor, ir, density, dt)
    !! Replaces main.f90:630-663 and rotor(ir)%calc_secAlpha() / %calc_force(density, dt) (forceCalcSwitch 0) down to the
    !! blade sums; the driver then calls rotor(ir)%sumBladeToNetForces() (classdef.f90:4954-4988).
    type(rotor_class), intent(inout) :: rotor(:)
    integer, intent(in) :: ir
    real(dp), intent(in) :: density, dt
    integer :: jr, ib, ns, k
    real(c_double), allocatable :: sec(:), loads(:), buf(:)
    ns = rotor(ir)%ns
    do jr = 1, size(rotor)                ! the bound vortices of every rotor are sources
      call gpu_sync_rotor(rotor(jr), jr, .false.)
    enddo
    allocate (sec(10*ns + 6), loads(12 + 25*ns))
    do ib = 1, rotor(ir)%nb
      associate (b => rotor(ir)%blade(ib))
        sec(1:3*ns) = reshape(b%secTauCapChord, [3*ns])
        sec(3*ns + 1:6*ns) = reshape(b%secNormalVec, [3*ns])
        sec(6*ns + 1:9*ns) = reshape(b%secCP, [3*ns])
        sec(9*ns + 1:10*ns) = b%secArea
        sec(10*ns + 1:10*ns + 3) = b%yAxisAziFlap
        sec(10*ns + 4:10*ns + 6) = b%zAxisAziFlap
      end associate
      call check(vlc_rotor_put_sections(ctx, ir - 1, ib - 1, sec))
    enddo
    call check(vlc_rotor_calc_velCPTotal(ctx, ir - 1))
    call check(vlc_rotor_calc_force(ctx, ir - 1, density, dt, rotor(ir)%Omega, rotor(ir)%spanwiseLiftSwitch))
    do ib = 1, rotor(ir)%nb
      if (ib > rotor(ir)%nbConvect .and. rotor(ir)%axisymmetrySwitch /= 1) cycle
      associate (b => rotor(ir)%blade(ib))
        allocate (buf(104*rotor(ir)%nc*ns))
        call check(vlc_rotor_get_wing(ctx, ir - 1, ib - 1, buf))
        b%wiP = reshape(transfer(buf, b%wiP), shape(b%wiP))   ! gamPrev, gamTrapz, delP, normalForce, velCPTotal ...
        deallocate (buf)
        call check(vlc_rotor_get_loads(ctx, ir - 1, ib - 1, loads))
        b%forceInertial = loads(1:3)
        b%lift = loads(4:6)
        b%drag = loads(7:9)
        b%liftUnsteady = loads(10:12)
        k = 12
        if (ib <= rotor(ir)%nbConvect) b%secChordwiseResVel = reshape(loads(k + 1:k + 3*ns), [3, ns])
        b%secDragDir = reshape(loads(k + 3*ns + 1:k + 6*ns), [3, ns])
        b%secLiftDir = reshape(loads(k + 6*ns + 1:k + 9*ns), [3, ns])
        b%secForceInertial = reshape(loads(k + 9*ns + 1:k + 12*ns), [3, ns])
        b%secLift = reshape(loads(k + 12*ns + 1:k + 15*ns), [3, ns])
        b%secDrag = reshape(loads(k + 15*ns + 1:k + 18*ns), [3, ns])
        b%secLiftUnsteady = reshape(loads(k + 18*ns + 1:k + 21*ns), [3, ns])
        k = 12 + 21*ns
        b%secAlpha = loads(k + 1:k + ns)
        b%secCL = loads(k + ns + 1:k + 2*ns)
        b%secCD = loads(k + 2*ns + 1:k + 3*ns)
        b%secCLu = loads(k + 3*ns + 1:k + 4*ns)
      end associate
    enddo
  end subroutine gpu_cp_forces

  subroutine gpu_download_wake(rotor)
    !! Bring the device's wake back into the driver's derived types (before wake plots / restart files): records of the
    !! current and the predicted wake, the prescribed helix, the velocity arrays of the convection driver.
    type(rotor_class), intent(inout) :: rotor(:)
    integer :: ir, ib
    real(c_double), allocatable :: buf(:)
    real(c_double) :: helix(2)
    do ir = 1, size(rotor)
      do ib = 1, rotor(ir)%nb
        if (rotor(ir)%nNwake > 0) then
          allocate (buf(50*size(rotor(ir)%blade(ib)%waN)))
          call check(vlc_rotor_get_nwake(ctx, ir - 1, ib - 1, 0_c_int, buf))
          rotor(ir)%blade(ib)%waN = reshape(transfer(buf, rotor(ir)%blade(ib)%waN), shape(rotor(ir)%blade(ib)%waN))
          deallocate (buf)
        endif
        if (rotor(ir)%nFwake > 0) then
          allocate (buf(13*size(rotor(ir)%blade(ib)%waF)))
          call check(vlc_rotor_get_fwake(ctx, ir - 1, ib - 1, 0_c_int, buf))
          rotor(ir)%blade(ib)%waF = transfer(buf, rotor(ir)%blade(ib)%waF)
          deallocate (buf)
        endif
        if (rotor(ir)%prescWakeNt > 0) then   ! records and fit parameters of the prescribed far wake made on the device
          allocate (buf(13*240))
          call check(vlc_rotor_get_pfwake(ctx, ir - 1, ib - 1, 0_c_int, buf, helix))
          rotor(ir)%blade(ib)%wapF%waF = transfer(buf, rotor(ir)%blade(ib)%wapF%waF)
          rotor(ir)%blade(ib)%wapF%helixPitch = helix(1)
          rotor(ir)%blade(ib)%wapF%helixRadius = helix(2)
          rotor(ir)%blade(ib)%wapF%isPresent = .true.
          deallocate (buf)
        endif
        if (rotor(ir)%nNwake > 0) then   ! the predicted records and the velocity arrays of the fdScheme in use (restart files)
          if (allocated(rotor(ir)%blade(ib)%waNPredicted)) then
            allocate (buf(50*size(rotor(ir)%blade(ib)%waNPredicted)))
            call check(vlc_rotor_get_nwake(ctx, ir - 1, ib - 1, 1_c_int, buf))
            rotor(ir)%blade(ib)%waNPredicted = reshape(transfer(buf, rotor(ir)%blade(ib)%waNPredicted), &
              & shape(rotor(ir)%blade(ib)%waNPredicted))
            deallocate (buf)
          endif
          if (allocated(rotor(ir)%blade(ib)%waFPredicted)) then
            if (size(rotor(ir)%blade(ib)%waFPredicted) > 0) then
              allocate (buf(13*size(rotor(ir)%blade(ib)%waFPredicted)))
              call check(vlc_rotor_get_fwake(ctx, ir - 1, ib - 1, 1_c_int, buf))
              rotor(ir)%blade(ib)%waFPredicted = transfer(buf, rotor(ir)%blade(ib)%waFPredicted)
              deallocate (buf)
            endif
          endif
          call get_vel(ir, ib, ARR_VEL, rotor(ir)%blade(ib)%velNwake, rotor(ir)%blade(ib)%velFwake)
          call get_vel(ir, ib, ARR_VEL1, rotor(ir)%blade(ib)%velNwake1, rotor(ir)%blade(ib)%velFwake1)
          call get_vel(ir, ib, ARR_PREDICTED, rotor(ir)%blade(ib)%velNwakePredicted, rotor(ir)%blade(ib)%velFwakePredicted)
          call get_vel(ir, ib, ARR_STEP, rotor(ir)%blade(ib)%velNwakeStep, rotor(ir)%blade(ib)%velFwakeStep)
          call get_vel(ir, ib, ARR_VEL2, rotor(ir)%blade(ib)%velNwake2, rotor(ir)%blade(ib)%velFwake2)
          call get_vel(ir, ib, ARR_VEL3, rotor(ir)%blade(ib)%velNwake3, rotor(ir)%blade(ib)%velFwake3)
        endif
      enddo
    enddo
  end subroutine gpu_download_wake

  subroutine get_vel(ir, ib, which, velN, velF)
    !! One pair of the convection driver's velocity arrays (classdef.f90:285-292), when this fdScheme allocated it
    integer, intent(in) :: ir, ib
    integer(c_int), intent(in) :: which
    real(dp), allocatable, intent(inout) :: velN(:, :, :), velF(:, :)
    real(c_double) :: none(1)
    if (.not. allocated(velN)) return
    if (allocated(velF)) then
      if (size(velF) > 0) then
        call check(vlc_rotor_get_wakevel(ctx, ir - 1, ib - 1, which, velN, velF))
        return
      endif
    endif
    call check(vlc_rotor_get_wakevel(ctx, ir - 1, ib - 1, which, velN, none))   ! nFwake = 0: nothing is written to it
  end subroutine get_vel

  subroutine put_vel(ir, ib, which, velN, velF)
    !! Upload twin of get_vel
    integer, intent(in) :: ir, ib
    integer(c_int), intent(in) :: which
    real(dp), allocatable, intent(in) :: velN(:, :, :), velF(:, :)
    real(c_double) :: none(1)
    if (.not. allocated(velN)) return
    if (allocated(velF)) then
      if (size(velF) > 0) then
        call check(vlc_rotor_put_wakevel(ctx, ir - 1, ib - 1, which, velN, velF))
        return
      endif
    endif
    none = 0._c_double
    call check(vlc_rotor_put_wakevel(ctx, ir - 1, ib - 1, which, velN, none))   ! nFwake = 0: nothing is read from it
  end subroutine put_vel

end module libGPU
