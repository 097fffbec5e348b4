! ISO_C_BINDING interfaces of librrnn_b200.so (include/rrnn.h).
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE: no Fortran compiler exists here (SURVEY.md section 0, F1).  The file is kept
! deliberately declarative -- one interface block per C entry point, scalars by value, arrays by reference as
! type(c_ptr) device addresses or contiguous real(c_float) host arrays -- so that a maintainer with gfortran /
! nvfortran can build it with `make -C fortran` (guarded on `command -v gfortran`).  Every compute entry point is
! exercised through the same C ABI from Python (tests/test_parity_gpu.py); nothing in the shim can change results.
module mo_rrnn_c_binding
  use, intrinsic :: iso_c_binding
  implicit none
  public

  ! rrnn_gas_t (include/rrnn.h): one gas of ty_gas_concs
  type, bind(C) :: rrnn_gas_t
    character(kind=c_char) :: name(32)
    type(c_ptr)            :: conc
    real(c_float)          :: value
    integer(c_int)         :: ndims
  end type rrnn_gas_t

  interface
    function rrnn_last_error() bind(C, name="rrnn_last_error") result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function rrnn_ctx_create(device, stream, ctx) bind(C, name="rrnn_ctx_create") result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr),    value :: stream
      type(c_ptr)           :: ctx
      integer(c_int)        :: rc
    end function
    function rrnn_ctx_destroy(ctx) bind(C, name="rrnn_ctx_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int)     :: rc
    end function
    function rrnn_ctx_set_flag(ctx, name, val) bind(C, name="rrnn_ctx_set_flag") result(rc)
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: ctx
      character(kind=c_char) :: name(*)
      integer(c_int), value :: val
      integer(c_int)     :: rc
    end function
    ! rrtmgp_network_type%load_netcdf  (neural/mod_network_rrtmgp.F90:58-122)
    function rrnn_model_load_netcdf(ctx, filename, model) bind(C, name="rrnn_model_load_netcdf") result(rc)
      import :: c_int, c_ptr, c_char
      type(c_ptr), value     :: ctx
      character(kind=c_char) :: filename(*)
      type(c_ptr)            :: model
      integer(c_int)         :: rc
    end function
    function rrnn_model_destroy(model) bind(C, name="rrnn_model_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: model
      integer(c_int)     :: rc
    end function
    ! ty_gas_optics_rrtmgp%load subset used by the NN path  (rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326)
    function rrnn_kdist_create(ctx, nbnd, ngpt, band_lims_gpt, ntemp, totplnk, temp_ref_min, totplnk_delta, &
                               solar_source, kd) bind(C, name="rrnn_kdist_create") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),    value :: ctx
      integer(c_int), value :: nbnd, ngpt, ntemp
      integer(c_int)        :: band_lims_gpt(2, *)
      type(c_ptr),    value :: totplnk, solar_source     ! host real(c_float) arrays or c_null_ptr
      real(c_float),  value :: temp_ref_min, totplnk_delta
      type(c_ptr)           :: kd
      integer(c_int)        :: rc
    end function
    function rrnn_kdist_set_tsi(kd, tsi) bind(C, name="rrnn_kdist_set_tsi") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),   value :: kd
      real(c_float), value :: tsi
      integer(c_int)       :: rc
    end function
    ! load_ext's solar tables (rrtmgp/mo_gas_optics_rrtmgp.F90:1317-1325) and set_solar_variability (:1058-1095)
    function rrnn_kdist_set_solar_tables(kd, solar_quiet, solar_facular, solar_sunspot) &
                                         bind(C, name="rrnn_kdist_set_solar_tables") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),   value :: kd
      real(c_float)        :: solar_quiet(*), solar_facular(*), solar_sunspot(*)
      integer(c_int)       :: rc
    end function
    function rrnn_kdist_set_solar_variability(kd, mg_index, sb_index, have_tsi, tsi) &
                                              bind(C, name="rrnn_kdist_set_solar_variability") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),    value :: kd
      real(c_float),  value :: mg_index, sb_index, tsi
      integer(c_int), value :: have_tsi
      integer(c_int)        :: rc
    end function
    function rrnn_kdist_get_solar_source(kd, solar_source) bind(C, name="rrnn_kdist_get_solar_source") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),   value :: kd
      real(c_float)        :: solar_source(*)
      integer(c_int)       :: rc
    end function
    ! optimal_angle_fit(2,nbnd) of load (:1163, 1210) and compute_optimal_angles (:1712-1758); device pointers
    function rrnn_kdist_set_optimal_angle_fit(kd, optimal_angle_fit) bind(C, name="rrnn_kdist_set_optimal_angle_fit") result(rc)
      import :: c_int, c_ptr, c_float
      type(c_ptr),   value :: kd
      real(c_float)        :: optimal_angle_fit(2,*)
      integer(c_int)       :: rc
    end function
    function rrnn_compute_optimal_angles(ctx, kd, nlay, ncol, tau, optimal_angles) &
                                         bind(C, name="rrnn_compute_optimal_angles") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd, tau, optimal_angles
      integer(c_int), value :: nlay, ncol
      integer(c_int)        :: rc
    end function
    ! ty_fluxes_byband%reduce (extensions/mo_fluxes_byband.F90:41-131): sum_byband, net_byband_full, net_*_precalc
    function rrnn_sum_byband(ctx, kd, nlev, ncol, gpt_flux, bnd_flux) bind(C, name="rrnn_sum_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd, gpt_flux, bnd_flux
      integer(c_int), value :: nlev, ncol
      integer(c_int)        :: rc
    end function
    function rrnn_net_byband(ctx, kd, nlev, ncol, gpt_flux_dn, gpt_flux_up, bnd_flux_net) &
                             bind(C, name="rrnn_net_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd, gpt_flux_dn, gpt_flux_up, bnd_flux_net
      integer(c_int), value :: nlev, ncol
      integer(c_int)        :: rc
    end function
    function rrnn_net_flux(ctx, n, flux_dn, flux_up, flux_net) bind(C, name="rrnn_net_flux") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr),       value :: ctx, flux_dn, flux_up, flux_net
      integer(c_size_t), value :: n
      integer(c_int)           :: rc
    end function
    ! gas_optics (LW), neural_nets present  (rrtmgp/mo_gas_optics_rrtmgp.F90:239-243, NN branch :368-411)
    function rrnn_gas_optics_lw(ctx, kd, models, nmodels, ncol, nlay, play, plev, tlay, tsfc, gases, ngas, tlev, &
                                tau, lay_source, lev_source, sfc_source, sfc_source_Jac) &
                                bind(C, name="rrnn_gas_optics_lw") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr),    value :: ctx, kd
      type(c_ptr)           :: models(*)
      integer(c_int), value :: nmodels, ncol, nlay, ngas
      type(c_ptr),    value :: play, plev, tlay, tsfc, tlev           ! device pointers
      type(rrnn_gas_t)      :: gases(*)
      type(c_ptr),    value :: tau, lay_source, lev_source, sfc_source, sfc_source_Jac
      integer(c_int)        :: rc
    end function
    ! gas_optics (SW)  (rrtmgp/mo_gas_optics_rrtmgp.F90:433-437, NN branch :529-573)
    function rrnn_gas_optics_sw(ctx, kd, models, ncol, nlay, play, plev, tlay, gases, ngas, tau, ssa, g, toa_src) &
                                bind(C, name="rrnn_gas_optics_sw") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr),    value :: ctx, kd
      type(c_ptr)           :: models(*)
      integer(c_int), value :: ncol, nlay, ngas
      type(c_ptr),    value :: play, plev, tlay
      type(rrnn_gas_t)      :: gases(*)
      type(c_ptr),    value :: tau, ssa, g, toa_src
      integer(c_int)        :: rc
    end function
    ! rte_lw  (rte/mo_rte_lw.F90:60-64)
    function rrnn_rte_lw(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux, tau, lay_source, lev_source, &
                         sfc_source, sfc_emis, flux_up, flux_dn) bind(C, name="rrnn_rte_lw") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd
      integer(c_int), value :: nlay, ncol, top_at_1, n_gauss_angles
      type(c_ptr),    value :: inc_flux, tau, lay_source, lev_source, sfc_source, sfc_emis, flux_up, flux_dn
      integer(c_int)        :: rc
    end function
    ! rte_sw  (rte/mo_rte_sw.F90:48-52)
    function rrnn_rte_sw(ctx, ngpt, nlay, ncol, top_at_1, mu0, inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt, inc_flux_dif, &
                         tau, ssa, g, flux_up, flux_dn, flux_dn_dir) bind(C, name="rrnn_rte_sw") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx
      integer(c_int), value :: ngpt, nlay, ncol, top_at_1
      type(c_ptr),    value :: mu0, inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt, inc_flux_dif, tau, ssa, g
      type(c_ptr),    value :: flux_up, flux_dn, flux_dn_dir
      integer(c_int)        :: rc
    end function
    ! whole block-loop body with HOST arrays (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446)
    function rrnn_lw_fluxes_host(ctx, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play, plev, tlay, tlev, &
                                 tsfc, sfc_emis, gases, ngas, flux_up, flux_dn) bind(C, name="rrnn_lw_fluxes_host") result(rc)
      import :: c_int, c_ptr, c_float, rrnn_gas_t
      type(c_ptr),    value :: ctx, kd
      type(c_ptr)           :: models(*)
      integer(c_int), value :: nmodels, ncol, nlay, top_at_1, n_gauss_angles, ngas
      real(c_float)         :: play(nlay, *), plev(nlay+1, *), tlay(nlay, *), tlev(nlay+1, *), tsfc(*), sfc_emis(*)
      type(rrnn_gas_t)      :: gases(*)
      real(c_float)         :: flux_up(nlay+1, *), flux_dn(nlay+1, *)
      integer(c_int)        :: rc
    end function
    function rrnn_sw_fluxes_host(ctx, kd, models, ncol, nlay, top_at_1, play, plev, tlay, mu0, sfc_alb, tsi, gases, ngas, &
                                 flux_up, flux_dn, flux_dn_dir) bind(C, name="rrnn_sw_fluxes_host") result(rc)
      import :: c_int, c_ptr, c_float, rrnn_gas_t
      type(c_ptr),    value :: ctx, kd
      type(c_ptr)           :: models(*)
      integer(c_int), value :: ncol, nlay, top_at_1, ngas
      real(c_float)         :: play(nlay, *), plev(nlay+1, *), tlay(nlay, *), mu0(*), sfc_alb(*)
      type(c_ptr),    value :: tsi                                  ! c_loc(tsi) or c_null_ptr
      type(rrnn_gas_t)      :: gases(*)
      real(c_float)         :: flux_up(nlay+1, *), flux_dn(nlay+1, *), flux_dn_dir(nlay+1, *)
      integer(c_int)        :: rc
    end function
    ! rte_lw for _2str clouds (re-scaled solution), lw_Ds, flux_up_Jac, g-point fluxes  (rte/mo_rte_lw.F90:60-64, 324-384);
    ! device pointers, c_null_ptr for absent optional arguments
    function rrnn_rte_lw_ext(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux, tau, ssa, g, lay_source, lev_source, &
                             sfc_source, sfc_emis, lw_Ds, sfc_source_Jac, flux_up, flux_dn, flux_up_Jac, gpt_flux_up, &
                             gpt_flux_dn) bind(C, name="rrnn_rte_lw_ext") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd
      integer(c_int), value :: nlay, ncol, top_at_1, n_gauss_angles
      type(c_ptr),    value :: inc_flux, tau, ssa, g, lay_source, lev_source, sfc_source, sfc_emis, lw_Ds, sfc_source_Jac
      type(c_ptr),    value :: flux_up, flux_dn, flux_up_Jac, gpt_flux_up, gpt_flux_dn
      integer(c_int)        :: rc
    end function
    ! rte_lw(..., use_2stream = .true.)  (rte/mo_rte_lw.F90:346-361, lw_solver_2stream)
    function rrnn_rte_lw_2stream(ctx, kd, nlay, ncol, top_at_1, inc_flux, tau, ssa, g, lev_source, sfc_source, sfc_emis, &
                                 flux_up, flux_dn, gpt_flux_up, gpt_flux_dn) bind(C, name="rrnn_rte_lw_2stream") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd
      integer(c_int), value :: nlay, ncol, top_at_1
      type(c_ptr),    value :: inc_flux, tau, ssa, g, lev_source, sfc_source, sfc_emis, flux_up, flux_dn, gpt_flux_up, gpt_flux_dn
      integer(c_int)        :: rc
    end function
    ! sw_solver_2stream with g-point fluxes  (rte/kernels/mo_rte_solver_kernels.F90:541-546)
    function rrnn_sw_solver_2stream_ext(ctx, ngpt, nlay, ncol, top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, sfc_alb_dir, &
                                        sfc_alb_dif, flux_up, flux_dn, flux_dir, gpt_flux_up, gpt_flux_dn, gpt_flux_dir) &
                                        bind(C, name="rrnn_sw_solver_2stream_ext") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx
      integer(c_int), value :: ngpt, nlay, ncol, top_at_1
      type(c_ptr),    value :: inc_flux, inc_flux_dif, tau, ssa, g, mu0, sfc_alb_dir, sfc_alb_dif
      type(c_ptr),    value :: flux_up, flux_dn, flux_dir, gpt_flux_up, gpt_flux_dn, gpt_flux_dir
      integer(c_int)        :: rc
    end function
    ! cloud_optics (LUT or Pade handle)  (extensions/cloud_optics/mo_cloud_optics.F90:354-535)
    function rrnn_cloud_optics(ctx, lut, ncol, nlay, clwp, ciwp, reliq, reice, tau, ssa, g) bind(C, name="rrnn_cloud_optics") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, lut
      integer(c_int), value :: ncol, nlay
      type(c_ptr),    value :: clwp, ciwp, reliq, reice, tau, ssa, g
      integer(c_int)        :: rc
    end function
    ! sampled_mask_max_ran / sampled_mask_exp_ran (overlap_param = c_null_ptr: maximum-random), draw_samples
    ! (extensions/cloud_optics/mo_cloud_sampling.F90:38-286); the mask is one byte per element, (ngpt,nlay,ncol)
    function rrnn_sampled_mask(ctx, ngpt, nlay, ncol, randoms, cloud_frac, overlap_param, cloud_mask) &
                               bind(C, name="rrnn_sampled_mask") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx
      integer(c_int), value :: ngpt, nlay, ncol
      type(c_ptr),    value :: randoms, cloud_frac, overlap_param, cloud_mask
      integer(c_int)        :: rc
    end function
    function rrnn_draw_samples(ctx, kd, nlay, ncol, cloud_mask, tau_bnd, ssa_bnd, g_bnd, tau_gpt, ssa_gpt, g_gpt) &
                               bind(C, name="rrnn_draw_samples") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx, kd
      integer(c_int), value :: nlay, ncol
      type(c_ptr),    value :: cloud_mask, tau_bnd, ssa_bnd, g_bnd, tau_gpt, ssa_gpt, g_gpt
      integer(c_int)        :: rc
    end function
    ! compute_heating_rate  (extensions/mo_heating_rates.F90:26-54)
    function rrnn_heating_rate(ctx, ncol, nlay, flux_up, flux_dn, plev, heating_rate) bind(C, name="rrnn_heating_rate") result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: ctx
      integer(c_int), value :: ncol, nlay
      type(c_ptr),    value :: flux_up, flux_dn, plev, heating_rate
      integer(c_int)        :: rc
    end function
  end interface

contains

  ! copy the thread-local C error string into the reference's character(len=128) error_msg convention
  function rrnn_error_msg(rc) result(error_msg)
    integer(c_int), intent(in) :: rc
    character(len=128)         :: error_msg
    character(kind=c_char), pointer :: s(:)
    integer :: i
    error_msg = ""
    if (rc == 0) return
    call c_f_pointer(rrnn_last_error(), s, [128])
    do i = 1, 128
      if (s(i) == c_null_char) exit
      error_msg(i:i) = s(i)
    end do
  end function rrnn_error_msg

end module mo_rrnn_c_binding
