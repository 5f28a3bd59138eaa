!! Additions to src/python/f90/python_particles.f90 when built with -DATX_GPU: particles_t keeps its
!! host arrays (Abox, Bbox, pbc, r_non_cyc, Z, el, revision counters) and mirrors them to the device
!! object that neighbors_t%atx_p points to.  Nothing else in the module changes.

  !! end of particles_set_cell (:286-346), after Bbox = Abox^-1 (gaussn) and I_changed_cell
  subroutine particles_set_cell_gpu(this, atx_p, ierror)
    use atx_c_api
    implicit none
    type(particles_t), intent(in)    :: this
    type(c_ptr),       intent(in)    :: atx_p
    integer, optional, intent(inout) :: ierror
    integer(c_int) :: ierr
    ! Abox(3,3), Bbox(3,3) column-major, pbc(3) as 0 / 1 -- exactly the Fortran fields
    ierr = atx_particles_set_cell(atx_p, this%Abox, this%Bbox, int(this%pbc, c_int))
    call atx_pass_error(ierr, ierror)
  endsubroutine particles_set_cell_gpu

  !! end of particles_update_elements (:617-658): el(1:nat) are the compact element ids
  subroutine particles_update_elements_gpu(this, atx_p, ierror)
    use atx_c_api
    implicit none
    type(particles_t), intent(in)    :: this
    type(c_ptr),       intent(in)    :: atx_p
    integer, optional, intent(inout) :: ierror
    integer(c_int) :: ierr
    ierr = atx_particles_set_elements(atx_p, int(this%nat, c_int), this%el)
    call atx_pass_error(ierr, ierror)
  endsubroutine particles_update_elements_gpu

  !! called from neighbors_update when pos_rev changed (:430-456): r_non_cyc(3, nat), unwrapped
  subroutine particles_sync_positions_gpu(this, atx_p, ierror)
    use atx_c_api
    implicit none
    type(particles_t), intent(in)    :: this
    type(c_ptr),       intent(in)    :: atx_p
    integer, optional, intent(inout) :: ierror
    integer(c_int) :: ierr
    ierr = atx_particles_set_positions(atx_p, int(this%nat, c_int), this%r_non_cyc)
    call atx_pass_error(ierr, ierror)
  endsubroutine particles_sync_positions_gpu

  !! error hand-over used by every shim: C ABI message -> the Fortran error stack (error.f90:177-209)
  subroutine atx_pass_error(ierr, ierror)
    use atx_c_api
    implicit none
    integer(c_int),    intent(in)    :: ierr
    integer, optional, intent(inout) :: ierror
    character(kind=c_char) :: buf(2048)
    character(2048)        :: msg
    integer                :: i, n
    if (ierr == 0) return
    n = atx_last_error(buf, 2048_c_int)
    msg = ""
    do i = 1, 2047
       if (buf(i) == C_NULL_CHAR) exit
       msg(i:i) = buf(i)
    enddo
    RAISE_ERROR(trim(msg), ierror)
  endsubroutine atx_pass_error
