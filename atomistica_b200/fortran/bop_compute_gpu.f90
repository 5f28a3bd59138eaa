!! Replacement for src/potentials/bop/default_compute_func.f90 (COMPUTE_FUNC) when built with
!! -DATX_GPU.  BOP_TYPE gains:  type(c_ptr) :: atx_pot = C_NULL_PTR
!! BIND_TO_FUNC (default_bind_to_func.f90) additionally calls, after its own bookkeeping:
!!     call atx_fill_params(this, par)                       ! copies this%db into type(atx_bop_params_t)
!!     ierr = atx_bop_create(atx_ctx, c_loc(par), this%atx_pot)
!!     ierr = atx_bop_bind_to(this%atx_pot, nl%atx_p, nl%atx_nl, p%nel, p%el2Z)
!! The same pattern applies to rebo2_module.f90:143-223 (atx_rebo2_*) and to
!! tabulated_alloy_eam.f90:360-415 (atx_eam_*, the spline arrays of simple_spline_t are passed
!! through type(atx_spline_t) with c_loc(this%fF(i)%y) etc.).

  subroutine COMPUTE_FUNC(this, p, nl, epot, f, wpot, mask, epot_per_at, &
       epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond, ierror)
    use atx_c_api
    implicit none
    type(BOP_TYPE),     intent(inout) :: this
    type(particles_t),  intent(inout) :: p
    type(neighbors_t),  intent(inout) :: nl
    real(DP),           intent(inout) :: epot
    real(DP),           intent(inout) :: f(3, p%maxnatloc)
    real(DP),           intent(inout) :: wpot(3, 3)
    integer,  optional, target, intent(in)    :: mask(p%maxnatloc)
    real(DP), optional, target, intent(inout) :: epot_per_at(p%maxnatloc)
    real(DP), optional, target, intent(inout) :: epot_per_bond(nl%neighbors_size)
    real(DP), optional, target, intent(inout) :: f_per_bond(3, nl%neighbors_size)
    real(DP), optional, target, intent(inout) :: wpot_per_at(3, 3, p%maxnatloc)
    real(DP), optional, target, intent(inout) :: wpot_per_bond(3, 3, nl%neighbors_size)
    integer,  optional, intent(out)   :: ierror

    type(c_ptr)    :: c_mask, c_epa, c_epb, c_fpb, c_wpa, c_wpb
    integer(c_int) :: ierr

    INIT_ERROR(ierror)
    call timer_start(BOP_NAME_STR // "_force")

    ! rebuilds on the GPU when pos_rev / cell_rev changed (neighbors_update, :430-456)
    call update(nl, p, ierror)
    PASS_ERROR(ierror)

    c_mask = C_NULL_PTR ; c_epa = C_NULL_PTR ; c_epb = C_NULL_PTR
    c_fpb  = C_NULL_PTR ; c_wpa = C_NULL_PTR ; c_wpb = C_NULL_PTR
    if (present(mask))           c_mask = c_loc(mask(1))
    if (present(epot_per_at))    c_epa  = c_loc(epot_per_at(1))
    if (present(epot_per_bond))  c_epb  = c_loc(epot_per_bond(1))
    if (present(f_per_bond))     c_fpb  = c_loc(f_per_bond(1, 1))
    if (present(wpot_per_at))    c_wpa  = c_loc(wpot_per_at(1, 1, 1))
    if (present(wpot_per_bond))  c_wpb  = c_loc(wpot_per_bond(1, 1, 1))

    ! energies, forces and virials are ADDED into the caller's arrays, as tls_reduce does
    ierr = atx_bop_energy_and_forces(this%atx_pot, nl%atx_p, nl%atx_nl, c_mask, epot, f, wpot, &
         c_epa, c_epb, c_fpb, c_wpa, c_wpb)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)

    call timer_stop(BOP_NAME_STR // "_force")
  endsubroutine COMPUTE_FUNC
