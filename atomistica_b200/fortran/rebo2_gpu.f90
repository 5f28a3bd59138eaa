!! Replacement bodies for src/potentials/bop/rebo2/rebo2_module.f90 when built with -DATX_GPU.
!! rebo2_t gains:  type(c_ptr) :: atx_pot = C_NULL_PTR
!! BIND_TO_FUNC (:70-135) keeps rebo2_db_init_with_parameters (rebo2_db.f90:81-303: constants, the
!! quintic g(cos theta) splines, table2d/table3d coefficient solves) on the host and then hands the
!! finished tables to the device; COMPUTE_FUNC (:143-223) becomes one call.

  !! end of BIND_TO_FUNC, after rebo2_db_init_with_parameters and the interaction range requests
  subroutine rebo2_bind_to_gpu(this, p, nl, ierror)
    use atx_c_api
    implicit none
    type(BOP_TYPE), target, intent(inout) :: this
    type(particles_t),      intent(inout) :: p
    type(neighbors_t),      intent(inout) :: nl
    integer,      optional, intent(inout) :: ierror

    type(atx_rebo2_params_t), target :: par       ! bind(C) image of atx_rebo2_params (atx_c_api.f90)
#ifdef SCREENING
    type(atx_rebo2_screening_t), target :: scr    ! rebo2_type.f90:59-71, 204-213
#endif
    integer(c_int) :: ierr

    ! scalars of Brenner 2002 Tables 2, 6, 7 and the derived constants of rebo2_db.f90:147-168
    par%cc_B1 = this%cc_B1 ; par%cc_B2 = this%cc_B2 ; par%cc_B3 = this%cc_B3
    par%cc_beta1 = this%cc_beta1 ; par%cc_beta2 = this%cc_beta2 ; par%cc_beta3 = this%cc_beta3
    par%cc_Q = this%cc_Q ; par%cc_A = this%cc_A ; par%cc_alpha = this%cc_alpha
    par%ch_B1 = this%ch_B1 ; par%ch_beta1 = this%ch_beta1 ; par%ch_Q = this%ch_Q
    par%ch_A = this%ch_A ; par%ch_alpha = this%ch_alpha
    par%hh_B1 = this%hh_B1 ; par%hh_beta1 = this%hh_beta1 ; par%hh_Q = this%hh_Q
    par%hh_A = this%hh_A ; par%hh_alpha = this%hh_alpha
    par%cc_g_theta = this%cc_g_theta
    par%cc_g1_coeff = reshape(this%cc_g1_coeff%c, (/ 18 /))     ! 3 intervals x 6 coefficients
    par%cc_g2_coeff = reshape(this%cc_g2_coeff%c, (/ 18 /))
    par%spgh = reshape(this%spgh, (/ 18 /)) ; par%igh = this%igh
    par%conalp = this%conalp ; par%conear = reshape(this%conear, (/ 36 /))
    par%conpe = this%conpe ; par%conan = this%conan ; par%conpf = this%conpf
    par%cut_in_l = this%cut_in_l ; par%cut_in_h = this%cut_in_h ; par%cut_in_h2 = this%cut_in_h2
    par%with_dihedral = merge(1, 0, this%with_dihedral)
    ! coefficient arrays of the bicubic / tricubic tables as table2d_init / table3d_init left them
    par%Fcc = c_loc(this%Fcc%coeff(1, 1, 1, 1)) ; par%Fch = c_loc(this%Fch%coeff(1, 1, 1, 1))
    par%Fhh = c_loc(this%Fhh%coeff(1, 1, 1, 1)) ; par%Tcc = c_loc(this%Tcc%coeff(1, 1, 1, 1))
    par%Pcc = c_loc(this%Pcc%coeff(1, 1, 1))    ; par%Pch = c_loc(this%Pch%coeff(1, 1, 1))

    if (c_associated(this%atx_pot)) ierr = atx_rebo2_destroy(this%atx_pot)
#ifdef SCREENING
    scr%cc_ar_r1 = this%cc_ar_r1 ; scr%cc_ar_r2 = this%cc_ar_r2
    scr%cc_bo_r1 = this%cc_bo_r1 ; scr%cc_bo_r2 = this%cc_bo_r2
    scr%cc_nc_r1 = this%cc_nc_r1 ; scr%cc_nc_r2 = this%cc_nc_r2
    scr%Cmin = this%Cmin ; scr%Cmax = this%Cmax
    ierr = atx_rebo2_create_screened(atx_ctx, c_loc(par), c_loc(scr), this%atx_pot)
#else
    ierr = atx_rebo2_create(atx_ctx, c_loc(par), this%atx_pot)
#endif
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
    ierr = atx_rebo2_bind_to(this%atx_pot, nl%atx_p, nl%atx_nl, int(p%nel, c_int), p%el2Z)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
  endsubroutine rebo2_bind_to_gpu


  subroutine COMPUTE_FUNC(this, p, nl, epot, f, wpot, epot_per_at, epot_per_bond, f_per_bond, &
       wpot_per_at, wpot_per_bond, ierror)
    use atx_c_api
    implicit none
    type(BOP_TYPE),     intent(inout) :: this
    type(particles_t),  intent(inout) :: p
    type(neighbors_t),  intent(inout) :: nl
    real(DP),           intent(inout) :: epot
    real(DP),           intent(inout) :: f(3, p%maxnatloc)
    real(DP),           intent(inout) :: wpot(3, 3)
    real(DP), optional, target, intent(inout) :: epot_per_at(p%maxnatloc)
    real(DP), optional, target, intent(inout) :: epot_per_bond(nl%neighbors_size)
    real(DP), optional, target, intent(inout) :: f_per_bond(3, nl%neighbors_size)
    real(DP), optional, target, intent(inout) :: wpot_per_at(3, 3, p%maxnatloc)
    real(DP), optional, target, intent(inout) :: wpot_per_bond(3, 3, nl%neighbors_size)
    integer,  optional, intent(out)   :: ierror

    type(c_ptr)    :: c_epa, c_epb, c_fpb, c_wpa, c_wpb
    integer(c_int) :: ierr

    INIT_ERROR(ierror)
    call timer_start(BOP_NAME_STR // "_force")
    call update(nl, p, ierror)
    PASS_ERROR(ierror)
    c_epa = C_NULL_PTR ; c_epb = C_NULL_PTR ; c_fpb = C_NULL_PTR ; c_wpa = C_NULL_PTR ; c_wpb = C_NULL_PTR
    if (present(epot_per_at))    c_epa = c_loc(epot_per_at(1))
    if (present(epot_per_bond))  c_epb = c_loc(epot_per_bond(1))
    if (present(f_per_bond))     c_fpb = c_loc(f_per_bond(1, 1))
    if (present(wpot_per_at))    c_wpa = c_loc(wpot_per_at(1, 1, 1))
    if (present(wpot_per_bond))  c_wpb = c_loc(wpot_per_bond(1, 1, 1))
    ! no mask: features per_at, per_bond only (rebo2.f90:22-27)
    ierr = atx_rebo2_energy_and_forces(this%atx_pot, nl%atx_p, nl%atx_nl, epot, f, wpot, &
         c_epa, c_epb, c_fpb, c_wpa, c_wpb)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
    call timer_stop(BOP_NAME_STR // "_force")
  endsubroutine COMPUTE_FUNC
