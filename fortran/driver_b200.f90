! Driver of klkuhlm/unconfined with its hot path moved to libunconfined_b200.so.
!
! What stays exactly as in the reference: reading the deck (read_input, driver_io.f90:30-666),
! the output header (driver_io.f90:668-845) and the layout of the output rows
! (driver.f90:234-273: screen average, RFMT/HFMT, t -> r -> z order).  What is replaced: the
! whole loop nest driver.f90:100-231 (de Hoog p-values, tanh-sinh and Gauss-Lobatto set-up,
! every lap_hank_soln call, level sums, extrapolation, Wynn-epsilon, both de Hoog inversions)
! by ONE call of unc_eval_grid through the ISO_C_BINDING module unconfined_b200.
!
! Build (not possible in this repository's image: it has no Fortran compiler; this file is
! syntax-reviewed only, the same argument lists are exercised through ctypes in tests/):
!   gfortran -O3 -cpp constants.f90 types.f90 utility.f90 driver_io.f90 \
!       unconfined_b200_mod.f90 driver_b200.f90 -L$REPO/unconfined_b200 -lunconfined_b200 \
!       -Wl,-rpath,$REPO/unconfined_b200 -o unconfined_b200
! invlap.f90, integration.f90, time.f90, laplace_hankel_solutions.f90 and cbessel.f90 are no
! longer on the link line (driver_io.f90 uses none of them).
!
! Command line / environment: none added.  UNC_FRESH_ABSCISSAE=1 in the environment asks for
! tanh-sinh abscissae recomputed for every (t,r) instead of the reference's behaviour
! (abscissae of the first (t,r) only, driver.f90:121-126,274).
program driver_b200
  use, intrinsic :: iso_c_binding
  use types
  use constants, only : DP, RFMT, HFMT
  use driver_io, only : read_input, write_timeseries_header, write_contour_header
  use unconfined_b200
  implicit none

  real(DP), parameter :: TEE_MULT = 2.0_DP      ! driver.f90:54
  integer, parameter :: OUT = 20                ! the unit read_input opens for the results

  type(invLaplace), target :: lap
  type(invHankel), target :: hank
  type(formation), target :: aq
  type(GaussLobatto) :: gl
  type(TanhSinh) :: ts
  type(well) :: pw
  type(solution) :: sol

  type(unc_params) :: prm
  real(c_double), allocatable, target :: sD(:,:,:), dsD(:,:,:), absc_scale(:,:)
  integer(c_int32_t), allocatable :: split(:), layer(:)
  type(c_ptr) :: scale_arg
  integer(c_int) :: rc
  integer :: it, ir, envlen
  character(len=8) :: envval

  call read_input(pw, aq, sol, lap, hank, gl, ts)

  ! sizes the header writers expect to find filled in (driver.f90:79-91)
  lap%np = 2*lap%M + 1
  ts%N = 2**ts%k - 1
  allocate(lap%p(lap%np), ts%kv(ts%R), ts%Nv(ts%R), ts%Q(ts%R), ts%hv(ts%R))
  ts%kv = [(ts%k - ts%R + it, it = 1, ts%R)]
  ts%Nv = 2**ts%kv - 1
  ts%hv = 4.0_DP/(2**ts%kv)

  if (sol%timeSeries) then
     call write_timeseries_header(pw, aq, sol, lap, hank, gl, ts, OUT)
  else
     call write_contour_header(pw, aq, sol, lap, hank, gl, ts, OUT)
  end if

  ! ---- the hot path: one call ------------------------------------------------------------
  call unc_fill_params(prm, pw, aq, sol, lap, hank, gl, ts, real(TEE_MULT, c_double))
  allocate(sD(sol%nz, sol%nr, sol%nt), dsD(sol%nz, sol%nr, sol%nt), &
       &   split(sol%nt), layer(sol%nz), absc_scale(sol%nr, sol%nt))
  split = int(hank%sv, c_int32_t)
  layer = int(sol%zLay, c_int32_t)

  ! reference-compatible: every (t,r) uses the abscissae of the first one; passing the scale also
  ! selects the reference's stale-infint carry (driver.f90:205-214) inside the library
  absc_scale = hank%j0z(hank%sv(1))/sol%rD(1)
  scale_arg = c_loc(absc_scale(1,1))
  call get_environment_variable('UNC_FRESH_ABSCISSAE', envval, envlen)
  if (envlen > 0) then
     if (envval(1:1) == '1') scale_arg = c_null_ptr
  end if

  rc = unc_eval_grid(prm, int(sol%nt, c_int32_t), sol%tD, split, int(sol%nr, c_int32_t), sol%rD, &
       & int(sol%nz, c_int32_t), sol%zD, layer, scale_arg, 0_c_int32_t, sD, dsD)
  if (rc /= UNC_OK) then
     write(*,'(A,I0,2A)') 'ERROR: unconfined_b200 returned ', rc, ': ', unc_last_error()
     stop 1
  end if

  ! ---- output, row layout of driver.f90:234-273 ---------------------------------------------
  do it = 1, sol%nt
     do ir = 1, sol%nr
        call write_rows(sD(:, ir, it), dsD(:, ir, it))
     end do
  end do
  rc = unc_shutdown()

contains

  subroutine write_rows(h, dh)
    real(c_double), intent(in) :: h(:), dh(:)
    real(DP) :: hobs, dhobs, scale
    integer :: m, nzo

    scale = 1.0_DP
    if (.not. sol%dimless) scale = sol%Hc

    if (sol%timeseries) then
       nzo = sol%zOrd
       if (.not. sol%piezometer .and. nzo > 1) then
          ! the reference's screen "trapezoid" exactly as it is written (driver.f90:236-239):
          ! the last point is counted once more and the divisor is 2*zOrd
          hobs = (h(1) + 2.0*sum(h(2:nzo)) + h(nzo))/(2*nzo)
          dhobs = (dh(1) + 2.0*sum(dh(2:nzo)) + dh(nzo))/(2*nzo)
       else
          hobs = h(1)
          dhobs = dh(1)
       end if
       if (sol%dimless) then
          write (OUT, '('//RFMT//',1X,2('//HFMT//',1X))') sol%tD(it), hobs, dhobs
       else
          write (OUT, '('//RFMT//',1X,2('//HFMT//',1X))') sol%t(it), hobs*scale, dhobs*scale
       end if
    else
       do m = 1, sol%nz
          if (sol%dimless) then
             write (OUT, '(2('//RFMT//',1X),2('//HFMT//',1X))') sol%zD(m), sol%rD(ir), h(m), dh(m)
          else
             write (OUT, '(2('//RFMT//',1X),2('//HFMT//',1X))') sol%z(m), sol%r(ir), h(m)*scale, dh(m)*scale
          end if
       end do
    end if
  end subroutine write_rows

end program driver_b200
