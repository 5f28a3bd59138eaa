!! Replacement bodies for src/potentials/eam/tabulated_alloy_eam.f90 when built with -DATX_GPU.
!! tabulated_alloy_eam_t gains:  type(c_ptr) :: atx_pot = C_NULL_PTR
!! (:360-415 energy_and_forces, :297-350 bind_to; init :147-259 keeps reading the setfl file and
!! building the simple_spline_t objects on the host -- their arrays are handed over unchanged).
!! Uses the interfaces of atx_c_api.f90 (generated from include/atomistica_b200.h).

  !! end of tabulated_alloy_eam_bind_to: after el2db has been filled (:318-340)
  subroutine tabulated_alloy_eam_bind_to_gpu(this, p, nl, ierror)
    use atx_c_api
    implicit none
    type(tabulated_alloy_eam_t), target, intent(inout) :: this
    type(particles_t),                   intent(inout) :: p
    type(neighbors_t),                   intent(inout) :: nl
    integer,                   optional, intent(inout) :: ierror

    ! atx_spline_t: bind(C) image of atx_spline, defined in atx_c_api.f90
    type(atx_spline_t), target :: fF(this%db%nel), frho(this%db%nel), fphi(this%db%nel, this%db%nel)
    integer(c_int) :: ierr
    integer        :: i, j

    do i = 1, this%db%nel
       call spline_image(this%fF(i), fF(i))
       call spline_image(this%frho(i), frho(i))
       do j = 1, this%db%nel
          call spline_image(this%fphi(i, j), fphi(i, j))      ! already scaled by 1/2 (:245)
       enddo
    enddo
    if (c_associated(this%atx_pot)) ierr = atx_eam_destroy(this%atx_pot)
    ierr = atx_eam_create(atx_ctx, int(this%db%nel, c_int), c_loc(fF), c_loc(frho), c_loc(fphi), &
         this%cutoff, this%atx_pot)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
    ! el2db(1:p%nel): database index of every particle element id, -1 when the filter excludes it
    ierr = atx_eam_bind_to(this%atx_pot, nl%atx_p, nl%atx_nl, int(p%nel, c_int), this%el2db)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)

  contains
    subroutine spline_image(s, o)
      type(simple_spline_t), target, intent(in)  :: s
      type(atx_spline_t),            intent(out) :: o
      o%n = s%n ; o%x0 = s%x0 ; o%dx = s%dx
      o%y = c_loc(s%y(1)) ; o%coeff1 = c_loc(s%coeff1(1)) ; o%coeff2 = c_loc(s%coeff2(1))
      o%coeff3 = c_loc(s%coeff3(1)) ; o%dcoeff1 = c_loc(s%dcoeff1(1)) ; o%dcoeff2 = c_loc(s%dcoeff2(1))
      o%dcoeff3 = c_loc(s%dcoeff3(1))
    endsubroutine
  endsubroutine tabulated_alloy_eam_bind_to_gpu


  subroutine tabulated_alloy_eam_energy_and_forces(this, p, nl, epot, f, wpot, mask, epot_per_at, &
       wpot_per_at, ierror)
    use atx_c_api
    implicit none
    type(tabulated_alloy_eam_t), intent(inout) :: this
    type(particles_t),           intent(in)    :: p
    type(neighbors_t),           intent(inout) :: nl
    real(DP),                    intent(inout) :: epot
    real(DP),                    intent(inout) :: f(3, p%nat)
    real(DP),                    intent(inout) :: wpot(3, 3)
    integer,  optional, target,  intent(in)    :: mask(p%nat)
    real(DP), optional, target,  intent(inout) :: epot_per_at(p%nat)
    real(DP), optional, target,  intent(inout) :: wpot_per_at(3, 3, p%nat)
    integer,  optional,          intent(inout) :: ierror

    type(c_ptr)    :: c_mask, c_epa, c_wpa
    integer(c_int) :: ierr

    call timer_start("tabulated_alloy_eam_energy_and_forces")
    INIT_ERROR(ierror)
    call update(nl, p, ierror)            ! rebuilds on the GPU (python_neighbors_gpu.f90)
    PASS_ERROR(ierror)
    c_mask = C_NULL_PTR ; c_epa = C_NULL_PTR ; c_wpa = C_NULL_PTR
    if (present(mask))         c_mask = c_loc(mask(1))
    if (present(epot_per_at))  c_epa  = c_loc(epot_per_at(1))
    if (present(wpot_per_at))  c_wpa  = c_loc(wpot_per_at(1, 1, 1))
    ! both passes (density, embedding + pair force) run on the device; results are ADDED
    ierr = atx_eam_energy_and_forces(this%atx_pot, nl%atx_p, nl%atx_nl, c_mask, epot, f, wpot, c_epa, c_wpa)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
    call timer_stop("tabulated_alloy_eam_energy_and_forces")
  endsubroutine tabulated_alloy_eam_energy_and_forces
