!! ISO_C_BINDING interface blocks for libatomistica_b200.so (include/atomistica_b200.h).
!! Drop this file into src/support/ of the reference tree; it has no dependencies.
!! NOT compiled in the build container (no Fortran compiler there) -- the same C entry points are
!! exercised through ctypes by tests/ with identical argument images.
module atx_c_api
  use, intrinsic :: iso_c_binding
  implicit none

  interface
     integer(c_int) function atx_ctx_create(device, ctx) bind(C, name="atx_ctx_create")
       import :: c_int, c_ptr
       integer(c_int), value :: device
       type(c_ptr)           :: ctx
     endfunction
     integer(c_int) function atx_last_error(buf, len) bind(C, name="atx_last_error")
       import :: c_int, c_char
       character(kind=c_char) :: buf(*)
       integer(c_int), value  :: len
     endfunction

     integer(c_int) function atx_particles_create(ctx, p) bind(C, name="atx_particles_create")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr)        :: p
     endfunction
     integer(c_int) function atx_particles_set_cell(p, Abox, Bbox, pbc) bind(C, name="atx_particles_set_cell")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value   :: p
       real(c_double)       :: Abox(3, 3), Bbox(3, 3)
       integer(c_int)       :: pbc(3)
     endfunction
     integer(c_int) function atx_particles_set_positions(p, nat, r) bind(C, name="atx_particles_set_positions")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: p
       integer(c_int), value :: nat
       real(c_double)        :: r(3, *)
     endfunction
     integer(c_int) function atx_particles_set_elements(p, nat, el) bind(C, name="atx_particles_set_elements")
       import :: c_int, c_ptr
       type(c_ptr), value    :: p
       integer(c_int), value :: nat
       integer(c_int)        :: el(*)
     endfunction

     integer(c_int) function atx_neighbors_create(ctx, avgn, nl) bind(C, name="atx_neighbors_create")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx
       integer(c_int), value :: avgn
       type(c_ptr)           :: nl
     endfunction
     integer(c_int) function atx_neighbors_request_interaction_range(nl, cutoff) &
          bind(C, name="atx_neighbors_request_interaction_range")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: nl
       real(c_double), value :: cutoff
     endfunction
     integer(c_int) function atx_neighbors_update(nl, p) bind(C, name="atx_neighbors_update")
       import :: c_int, c_ptr
       type(c_ptr), value :: nl, p
     endfunction
     integer(c_int) function atx_neighbors_get_info(nl, npairs, nebmax, n_cells, stencil) &
          bind(C, name="atx_neighbors_get_info")
       import :: c_int, c_ptr, c_long_long
       type(c_ptr), value     :: nl
       integer(c_long_long)   :: npairs
       integer(c_int)         :: nebmax, n_cells(3), stencil(3)
     endfunction
     integer(c_int) function atx_neighbors_copy_to_host(nl, seed, last, neighbors, dc, capacity) &
          bind(C, name="atx_neighbors_copy_to_host")
       import :: c_int, c_ptr, c_intptr_t, c_long_long
       type(c_ptr), value          :: nl
       integer(c_intptr_t)         :: seed(*), last(*)
       integer(c_int)              :: neighbors(*), dc(3, *)
       integer(c_long_long), value :: capacity
     endfunction

     integer(c_int) function atx_bop_create(ctx, par, pot) bind(C, name="atx_bop_create")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: par      ! c_loc of a type(atx_bop_params_t)
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_bop_bind_to(pot, p, nl, nel, el2Z) bind(C, name="atx_bop_bind_to")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot, p, nl
       integer(c_int), value :: nel
       integer(c_int)        :: el2Z(*)
     endfunction
     integer(c_int) function atx_bop_energy_and_forces(pot, p, nl, mask, epot, f, wpot, epot_per_at, &
          epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond) bind(C, name="atx_bop_energy_and_forces")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: pot, p, nl
       type(c_ptr), value :: mask            ! c_loc(mask) or C_NULL_PTR
       real(c_double)     :: epot, f(3, *), wpot(3, 3)
       type(c_ptr), value :: epot_per_at, epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond
     endfunction

     integer(c_int) function atx_eam_create(ctx, ndb, fF, frho, fphi, cutoff, pot) bind(C, name="atx_eam_create")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: ndb
       type(c_ptr), value    :: fF, frho, fphi   ! arrays of type(atx_spline_t)
       real(c_double), value :: cutoff
       type(c_ptr)           :: pot
     endfunction
     integer(c_int) function atx_eam_bind_to(pot, p, nl, nel, el2db) bind(C, name="atx_eam_bind_to")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot, p, nl
       integer(c_int), value :: nel
       integer(c_int)        :: el2db(*)
     endfunction
     integer(c_int) function atx_eam_energy_and_forces(pot, p, nl, mask, epot, f, wpot, epot_per_at, &
          wpot_per_at) bind(C, name="atx_eam_energy_and_forces")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: pot, p, nl, mask
       real(c_double)     :: epot, f(3, *), wpot(3, 3)
       type(c_ptr), value :: epot_per_at, wpot_per_at
     endfunction

     integer(c_int) function atx_rebo2_create(ctx, par, pot) bind(C, name="atx_rebo2_create")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, par
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_rebo2_create_screened(ctx, par, scr, pot) &
          bind(C, name="atx_rebo2_create_screened")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, par
       type(c_ptr), value :: scr      ! c_loc of a type(atx_rebo2_screening_t)
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_rebo2_bind_to(pot, p, nl, nel, el2Z) bind(C, name="atx_rebo2_bind_to")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot, p, nl
       integer(c_int), value :: nel
       integer(c_int)        :: el2Z(*)
     endfunction
     integer(c_int) function atx_rebo2_energy_and_forces(pot, p, nl, epot, f, wpot, epot_per_at, &
          epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond) bind(C, name="atx_rebo2_energy_and_forces")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: pot, p, nl
       real(c_double)     :: epot, f(3, *), wpot(3, 3)
       type(c_ptr), value :: epot_per_at, epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond
     endfunction

     ! --- screened variants, Juslin, funcfl EAM, pair styles, LAMMPS-style lists, output mode ---
     integer(c_int) function atx_bop_create_screened(ctx, par, scr, pot) bind(C, name="atx_bop_create_screened")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: par      ! c_loc of a type(atx_bop_params_t)
       type(c_ptr), value :: scr      ! c_loc of a type(atx_bop_screening_t): or1 or2 bor1 bor2 Cmin Cmax
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_bop_create_juslin(ctx, par, pot) bind(C, name="atx_bop_create_juslin")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: par      ! c_loc of a type(atx_juslin_params_t), database AFTER the mirroring of bind_to
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_eam_create_funcfl(ctx, fF, frho, fZ, cutoff, pot) bind(C, name="atx_eam_create_funcfl")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       type(c_ptr), value    :: fF, frho, fZ     ! c_loc of one type(atx_spline_t) each; fZ already scaled
       real(c_double), value :: cutoff
       type(c_ptr)           :: pot
     endfunction
     integer(c_int) function atx_pair_create(ctx, par, pot) bind(C, name="atx_pair_create")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: par      ! c_loc of a type(atx_pair_params_t)
       type(c_ptr)        :: pot
     endfunction
     integer(c_int) function atx_pair_bind_to(pot, p, nl, el1, el2) bind(C, name="atx_pair_bind_to")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot, p, nl
       integer(c_int), value :: el1, el2          ! this%el1, this%el2 of filter_from_string
     endfunction
     integer(c_int) function atx_pair_energy_and_forces(pot, p, nl, mask, epot, f, wpot, epot_per_at, &
          wpot_per_at) bind(C, name="atx_pair_energy_and_forces")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: pot, p, nl, mask
       real(c_double)     :: epot, f(3, *), wpot(3, 3)
       type(c_ptr), value :: epot_per_at, wpot_per_at
     endfunction
     integer(c_int) function atx_neighbors_set_external(nl, p, natloc, inum, ilist, numneigh, firstneigh) &
          bind(C, name="atx_neighbors_set_external")
       import :: c_int, c_ptr
       type(c_ptr), value    :: nl, p
       integer(c_int), value :: natloc, inum
       type(c_ptr), value    :: ilist, numneigh, firstneigh   ! LAMMPS NeighList arrays (int*, int*, int**)
     endfunction
     integer(c_int) function atx_neighbors_get_counters(nl, nbuilds, nreused) bind(C, name="atx_neighbors_get_counters")
       import :: c_int, c_ptr, c_long_long
       type(c_ptr), value   :: nl
       integer(c_long_long) :: nbuilds, nreused
     endfunction
     integer(c_int) function atx_bop_set_store_outputs(pot, on) bind(C, name="atx_bop_set_store_outputs")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot
       integer(c_int), value :: on
     endfunction
     integer(c_int) function atx_eam_set_store_outputs(pot, on) bind(C, name="atx_eam_set_store_outputs")
       import :: c_int, c_ptr
       type(c_ptr), value    :: pot
       integer(c_int), value :: on
     endfunction
  endinterface

  !> mirrors of atx_bop_screening, atx_juslin_params, atx_pair_params (include/atomistica_b200.h)
  type, bind(C) :: atx_bop_screening_t
     real(c_double) :: or1(6), or2(6), bor1(6), bor2(6), Cmin(6), Cmax(6)
  endtype atx_bop_screening_t
  !> rebo2_scr: cc_ar_r1 ... Cmax of rebo2_type.f90:204-213
  type, bind(C) :: atx_rebo2_screening_t
     real(c_double) :: cc_ar_r1, cc_ar_r2, cc_bo_r1, cc_bo_r2, cc_nc_r1, cc_nc_r2, Cmin, Cmax
  endtype atx_rebo2_screening_t
  type, bind(C) :: atx_juslin_params_t
     integer(c_int) :: nel, Z(3)
     real(c_double) :: D0(9), r0(9), S(9), beta(9), gamma(9), c(9), d(9), h(9), n(9), r1(9), r2(9)
     real(c_double) :: alpha(27), omega(27)
     integer(c_int) :: m(27)
  endtype atx_juslin_params_t
  type, bind(C) :: atx_pair_params_t
     integer(c_int) :: kind            ! 1 LJCut, 2 Harmonic, 3 DoubleHarmonic
     real(c_double) :: p(8)
     integer(c_int) :: shift
  endtype atx_pair_params_t

  !> mirror of atx_spline (include/atomistica_b200.h)
  type, bind(C) :: atx_spline_t
     integer(c_int) :: n
     real(c_double) :: x0, dx
     type(c_ptr)    :: y, coeff1, coeff2, coeff3, dcoeff1, dcoeff2, dcoeff3
  endtype atx_spline_t

  !> one process-wide context (device 0 unless ATOMISTICA_DEVICE is set by the host)
  type(c_ptr), save :: atx_ctx = C_NULL_PTR

contains

  !> push the library's message on the reference's error stack (src/support/error.f90)
  subroutine atx_pass_error(ierr, error)
    use error_module
    integer(c_int), intent(in)       :: ierr
    integer, optional, intent(inout) :: error
    character(kind=c_char)           :: buf(2048)
    character(2048)                  :: msg
    integer                          :: i, e
    if (ierr == 0) return
    e = atx_last_error(buf, 2048)
    msg = ""
    do i = 1, 2048
       if (buf(i) == C_NULL_CHAR) exit
       msg(i:i) = buf(i)
    enddo
    call push_error_with_info(trim(msg), "libatomistica_b200", 0, ERROR_UNSPECIFIED)
    if (present(error)) then
       error = ERROR_UNSPECIFIED
    else
       call error_abort(ERROR_UNSPECIFIED)
    endif
  endsubroutine atx_pass_error

endmodule atx_c_api
