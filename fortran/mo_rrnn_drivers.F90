! Reference-facing veneer: the block-loop body of examples/rfmip-clear-sky/rrtmgp_rfmip_{lw,sw}.F90 as two calls that
! keep the reference's argument meaning and character(len=128) error convention.  NOT COMPILED HERE (see
! mo_rrnn_c_binding.F90).  A host model that holds p_lay/t_lay/... as real(sp) (nlay,ncol) arrays replaces
!     k_dist%gas_optics(p_lay, p_lev, t_lay, sfc_t, gas_concs, optical_props, source, tlev=t_lev, neural_nets=nets)
!     rte_lw(optical_props, top_at_1, source, sfc_emis_spec, fluxes)
! by  error_msg = rrnn_lw(state, p_lay, p_lev, t_lay, t_lev, sfc_t, sfc_emis, gases, top_at_1, flux_up, flux_dn).
module mo_rrnn_drivers
  use, intrinsic :: iso_c_binding
  use mo_rrnn_c_binding
  implicit none
  private
  public :: ty_rrnn_state, rrnn_init_lw, rrnn_lw, rrnn_sw, rrnn_finalize

  type :: ty_rrnn_state
    type(c_ptr) :: ctx = c_null_ptr, kd = c_null_ptr
    type(c_ptr) :: models(2) = c_null_ptr
    integer     :: nmodels = 0
  end type

contains

  function rrnn_init_lw(this, device, tau_file, pfrac_file, band_lims_gpt, totplnk, temp_ref_min, totplnk_delta) result(error_msg)
    type(ty_rrnn_state), intent(inout) :: this
    integer,             intent(in)    :: device
    character(len=*),    intent(in)    :: tau_file, pfrac_file
    integer(c_int),      intent(in)    :: band_lims_gpt(:,:)          ! (2, nbnd)
    real(c_float), target, intent(in)  :: totplnk(:,:)                ! (nPlanckTemp, nbnd)
    real(c_float),       intent(in)    :: temp_ref_min, totplnk_delta
    character(len=128)                 :: error_msg
    integer(c_int) :: rc
    rc = rrnn_ctx_create(int(device, c_int), c_null_ptr, this%ctx)
    if (rc == 0) rc = rrnn_model_load_netcdf(this%ctx, trim(tau_file)//c_null_char, this%models(1))
    if (rc == 0) rc = rrnn_model_load_netcdf(this%ctx, trim(pfrac_file)//c_null_char, this%models(2))
    this%nmodels = 2
    if (rc == 0) rc = rrnn_kdist_create(this%ctx, int(size(band_lims_gpt, 2), c_int), int(maxval(band_lims_gpt), c_int), &
                                        band_lims_gpt, int(size(totplnk, 1), c_int), c_loc(totplnk), temp_ref_min, &
                                        totplnk_delta, c_null_ptr, this%kd)
    error_msg = rrnn_error_msg(rc)
  end function

  function rrnn_lw(this, play, plev, tlay, tlev, tsfc, sfc_emis, gases, top_at_1, flux_up, flux_dn) result(error_msg)
    type(ty_rrnn_state), intent(in)    :: this
    real(c_float),       intent(in)    :: play(:,:), plev(:,:), tlay(:,:), tlev(:,:), tsfc(:), sfc_emis(:)
    type(rrnn_gas_t),    intent(in)    :: gases(:)
    logical,             intent(in)    :: top_at_1
    real(c_float),       intent(out)   :: flux_up(:,:), flux_dn(:,:)
    character(len=128)                 :: error_msg
    error_msg = rrnn_error_msg(rrnn_lw_fluxes_host(this%ctx, this%kd, this%models, int(this%nmodels, c_int), &
                  int(size(play, 2), c_int), int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), 1_c_int, &
                  play, plev, tlay, tlev, tsfc, sfc_emis, gases, int(size(gases), c_int), flux_up, flux_dn))
  end function

  function rrnn_sw(this, play, plev, tlay, mu0, sfc_alb, gases, top_at_1, flux_up, flux_dn, flux_dn_dir) result(error_msg)
    type(ty_rrnn_state), intent(in)    :: this
    real(c_float),       intent(in)    :: play(:,:), plev(:,:), tlay(:,:), mu0(:), sfc_alb(:)
    type(rrnn_gas_t),    intent(in)    :: gases(:)
    logical,             intent(in)    :: top_at_1
    real(c_float),       intent(out)   :: flux_up(:,:), flux_dn(:,:), flux_dn_dir(:,:)
    character(len=128)                 :: error_msg
    error_msg = rrnn_error_msg(rrnn_sw_fluxes_host(this%ctx, this%kd, this%models, int(size(play, 2), c_int), &
                  int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), play, plev, tlay, mu0, sfc_alb, c_null_ptr, &
                  gases, int(size(gases), c_int), flux_up, flux_dn, flux_dn_dir))
  end function

  subroutine rrnn_finalize(this)
    type(ty_rrnn_state), intent(inout) :: this
    integer(c_int) :: rc, i
    do i = 1, this%nmodels
      rc = rrnn_model_destroy(this%models(i))
    end do
    rc = rrnn_ctx_destroy(this%ctx)
  end subroutine

end module mo_rrnn_drivers
