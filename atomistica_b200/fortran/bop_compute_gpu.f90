!! Replacement for src/potentials/bop/default_compute_func.f90 (COMPUTE_FUNC) and the tail of
!! default_bind_to_func.f90 (BIND_TO_FUNC) when built with -DATX_GPU.
!! BOP_TYPE gains:  type(c_ptr) :: atx_pot = C_NULL_PTR
!! Each module (tersoff.f90, kumagai.f90, brenner.f90 and their *_scr.f90 twins) defines, next to BOP_NAME,
!!     #define ATX_BOP_KIND  1        (the preprocessor needs the number: 1 = ATX_BOP_TERSOFF, 2 = ATX_BOP_KUMAGAI,
!!                                     3 = ATX_BOP_BRENNER of the header)
!! and BIND_TO_FUNC calls  bop_bind_to_gpu(this, p, nl, ierror)  after its own bookkeeping (:25-146).
!! Uses atx_c_api.f90 (interfaces, the bind(C) derived types and the ATX_* constants are generated from
!! include/atomistica_b200.h).

  !! end of BIND_TO_FUNC: this%db is complete, Z2db is built, the interaction ranges are requested
  subroutine bop_bind_to_gpu(this, p, nl, ierror)
    use atx_c_api
    implicit none
    type(BOP_TYPE),    target, intent(inout) :: this
    type(particles_t),         intent(inout) :: p
    type(neighbors_t),         intent(inout) :: nl
    integer,         optional, intent(inout) :: ierror

    type(atx_bop_params_t), target    :: par
#ifdef SCREENING
    type(atx_bop_screening_t), target :: scr
#endif
    integer(c_int) :: ierr
    integer        :: i, np

    par%kind = ATX_BOP_KIND
    par%nel  = int(this%db%nel, c_int)
    np       = this%db%nel*(this%db%nel+1)/2            ! PAIR_INDEX order, macros.inc:123
    par%Z    = 0
    do i = 1, this%db%nel
       par%Z(i) = int(atomic_number(a2s(this%db%el(:, i))), c_int)
    enddo
    par%r1(1:np) = this%db%r1(1:np) ; par%r2(1:np) = this%db%r2(1:np)
#if ATX_BOP_KIND == 1
    ! tersoff_params.f90:33-84
    par%A(1:np) = this%db%A(1:np) ; par%B(1:np) = this%db%B(1:np) ; par%xi(1:np) = this%db%xi(1:np)
    par%lambda(1:np) = this%db%lambda(1:np) ; par%mu(1:np) = this%db%mu(1:np)
    par%omega(1:np) = this%db%omega(1:np) ; par%mubo(1:np) = this%db%mubo(1:np)
    par%m(1:np) = int(this%db%m(1:np), c_int)
    do i = 1, this%db%nel
       par%beta(i) = this%db%beta(i) ; par%n(i) = this%db%n(i) ; par%c(i) = this%db%c(i)
       par%d(i) = this%db%d(i) ; par%h(i) = this%db%h(i)
    enddo
#elif ATX_BOP_KIND == 2
    ! kumagai_params.f90:30-118: lambda1, lambda2, alpha and the integer beta travel in lambda, mu, mubo, m
    par%A(1:np) = this%db%A(1:np) ; par%B(1:np) = this%db%B(1:np)
    par%lambda(1:np) = this%db%lambda1(1:np) ; par%mu(1:np) = this%db%lambda2(1:np)
    par%mubo(1:np) = this%db%alpha(1:np) ; par%m(1:np) = int(this%db%beta(1:np), c_int)
    do i = 1, this%db%nel
       par%eta(i) = this%db%eta(i) ; par%delta(i) = this%db%delta(i) ; par%h(i) = this%db%h(i)
       par%c1(i) = this%db%c1(i) ; par%c2(i) = this%db%c2(i) ; par%c3(i) = this%db%c3(i)
       par%c4(i) = this%db%c4(i) ; par%c5(i) = this%db%c5(i)
    enddo
#else
    ! brenner_params.f90:33-70: every parameter is pair-indexed; beta, c, d, h, n -> pbeta, pc, pd, ph, pn
    par%D0(1:np) = this%db%D0(1:np) ; par%r0(1:np) = this%db%r0(1:np) ; par%S(1:np) = this%db%S(1:np)
    par%pbeta(1:np) = this%db%beta(1:np) ; par%gamma(1:np) = this%db%gamma(1:np)
    par%pc(1:np) = this%db%c(1:np) ; par%pd(1:np) = this%db%d(1:np) ; par%ph(1:np) = this%db%h(1:np)
    par%mubo(1:np) = this%db%mu(1:np) ; par%pn(1:np) = this%db%n(1:np)
    par%m(1:np) = int(this%db%m(1:np), c_int)
#endif

    if (c_associated(this%atx_pot)) ierr = atx_bop_destroy(this%atx_pot)
#ifdef SCREENING
    scr%or1(1:np) = this%db%or1(1:np) ; scr%or2(1:np) = this%db%or2(1:np)
    scr%bor1(1:np) = this%db%bor1(1:np) ; scr%bor2(1:np) = this%db%bor2(1:np)
    scr%Cmin(1:np) = this%db%Cmin(1:np) ; scr%Cmax(1:np) = this%db%Cmax(1:np)
    ierr = atx_bop_create_screened(atx_ctx, c_loc(par), c_loc(scr), this%atx_pot)
#else
    ierr = atx_bop_create(atx_ctx, c_loc(par), this%atx_pot)
#endif
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
    ierr = atx_bop_bind_to(this%atx_pot, nl%atx_p, nl%atx_nl, int(p%nel, c_int), p%el2Z)
    call atx_pass_error(ierr, ierror)
    PASS_ERROR(ierror)
  endsubroutine bop_bind_to_gpu


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
