!! Replacement bodies for src/python/f90/python_neighbors.f90 (compile with -DATX_GPU).
!! neighbors_t gains two fields:   type(c_ptr) :: atx_nl = C_NULL_PTR, atx_p = C_NULL_PTR
!! refresh_neighbor_list keeps its bookkeeping (:459-564); only the two hot routines change.

  !> was: neighbors_binning_update (:904-959) + fill_neighbor_list (:570-754)
  recursive subroutine fill_neighbor_list(this, p, error)
    use atx_c_api
    implicit none
    type(neighbors_t), intent(inout)  :: this
    type(particles_t), intent(in)     :: p
    integer, optional, intent(inout)  :: error
    integer(c_int)        :: ierr, nebmax, ncells(3), sten(3)
    integer(c_long_long)  :: npairs

    INIT_ERROR(error)
    call timer_start("fill_neighbor_list")

    if (.not. c_associated(atx_ctx)) then
       ierr = atx_ctx_create(0_c_int, atx_ctx) ; call atx_pass_error(ierr, error) ; PASS_ERROR(error)
    endif
    if (.not. c_associated(this%atx_nl)) then
       ierr = atx_neighbors_create(atx_ctx, int(this%avgn, c_int), this%atx_nl)
       call atx_pass_error(ierr, error) ; PASS_ERROR(error)
       ierr = atx_particles_create(atx_ctx, this%atx_p)
       call atx_pass_error(ierr, error) ; PASS_ERROR(error)
       ierr = atx_neighbors_request_interaction_range(this%atx_nl, this%cutoff)
    endif

    ! device mirror of particles_t: cell, positions (unwrapped r_non_cyc), element ids
    ierr = atx_particles_set_cell(this%atx_p, p%Abox, p%Bbox, p%pbc)
    ierr = atx_particles_set_elements(this%atx_p, int(p%nat, c_int), p%el)
    ierr = atx_particles_set_positions(this%atx_p, int(p%nat, c_int), p%r_non_cyc)
    call atx_pass_error(ierr, error) ; PASS_ERROR(error)

    ! cell binning + pair search on the GPU; raises "Neighbor list overflow" like :716-718
    ierr = atx_neighbors_update(this%atx_nl, this%atx_p)
    call atx_pass_error(ierr, error) ; PASS_ERROR(error)

    ! host view for callers that read seed/last/neighbors/dc directly (neighbors_wrap.f90,
    ! pair potentials, coulomb): identical layout and ORDER as the CPU build
    ierr = atx_neighbors_copy_to_host(this%atx_nl, this%seed, this%last, this%neighbors, this%dc, &
         int(this%neighbors_size, c_long_long))
    call atx_pass_error(ierr, error) ; PASS_ERROR(error)

    ierr = atx_neighbors_get_info(this%atx_nl, npairs, nebmax, ncells, sten)
    this%nupdate = this%nupdate + 1
    this%avgnn   = this%avgnn + real(npairs, DP)/p%nat
    call timer_stop("fill_neighbor_list")
  endsubroutine fill_neighbor_list

  !> binning happens inside atx_neighbors_update
  subroutine neighbors_binning_update(this, p, error)
    implicit none
    type(neighbors_t), intent(inout)  :: this
    type(particles_t), intent(inout)  :: p
    integer, optional, intent(inout)  :: error
    INIT_ERROR(error)
  endsubroutine neighbors_binning_update
