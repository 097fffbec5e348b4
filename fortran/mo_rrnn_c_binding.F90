! ISO_C_BINDING interfaces of librrnn_b200.so -- GENERATED from include/rrnn.h by tools/gen_fortran_binding.py; do not edit.
!
! One interface block per C entry point (103 of 103).  Scalars by value; handles, device addresses and host
! arrays as type(c_ptr) values (host arrays: c_loc(a)); out-arguments and small integer / double arrays by reference.
! The comment above each block is the one the header carries: it cites the reference interface the entry point replaces.
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler, SURVEY.md section 0 F1): `make -C fortran` builds it where
! gfortran / nvfortran exist; tests/test_fortran_cpu.py parses it (numpy.f2py.crackfortran) and checks every bind(C) name,
! argument count and argument kind against the header.
module mo_rrnn_c_binding
  use, intrinsic :: iso_c_binding
  implicit none
  public

  ! rrnn_gas_t (include/rrnn.h): one gas of ty_gas_concs (rrtmgp/mo_gas_concentrations.F90:50-88)
  type, bind(C) :: rrnn_gas_t
    character(kind=c_char) :: name(32)
    type(c_ptr)            :: conc
    real(c_float)          :: value
    integer(c_int)         :: ndims
  end type rrnn_gas_t

  ! activation codes (neural/mod_layer.F90:64-95) and MLP kernel ids
  integer(c_int), parameter :: RRNN_ACT_LINEAR = 0, RRNN_ACT_SOFTSIGN = 1, RRNN_ACT_RELU = 2, RRNN_ACT_SIGMOID = 3, &
                               RRNN_ACT_HARD_SIGMOID = 4
  integer(c_int), parameter :: RRNN_NN_KERNEL_NONE = 0, RRNN_NN_KERNEL_FFMA = 1, RRNN_NN_KERNEL_TCGEN05 = 2

  interface
    ! ------------------------------------------------------------------------------------------------ library / context
    function rrnn_last_error() bind(C, name="rrnn_last_error") result(rc)
      import :: c_ptr
      type(c_ptr) :: rc
    end function rrnn_last_error
    function rrnn_version() bind(C, name="rrnn_version") result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function rrnn_version
    function rrnn_device_count() bind(C, name="rrnn_device_count") result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function rrnn_device_count
    ! stream: a cudaStream_t; NULL = the legacy default stream (the one PyTorch uses by default).
    function rrnn_ctx_create(device, stream, out_h) bind(C, name="rrnn_ctx_create") result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr), value :: stream
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_ctx_create
    function rrnn_ctx_destroy(ctx) bind(C, name="rrnn_ctx_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function rrnn_ctx_destroy
    function rrnn_ctx_set_stream(ctx, stream) bind(C, name="rrnn_ctx_set_stream") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: stream
      integer(c_int) :: rc
    end function rrnn_ctx_set_stream
    function rrnn_ctx_stream(ctx) bind(C, name="rrnn_ctx_stream") result(rc)
      import :: c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr) :: rc
    end function rrnn_ctx_stream
    function rrnn_ctx_synchronize(ctx) bind(C, name="rrnn_ctx_synchronize") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function rrnn_ctx_synchronize
    ! Run-time flags of rte/mo_rte_rrtmgp_config.F90:23-40. lw_source_bug_compat = 1 (default) reproduces lw_source_noscat
    ! ignoring top_at_1 (rte/kernels/mo_rte_solver_kernels.F90:770-773); 0 orients the level sources physically for top_at_1 =
    ! false. solver_wide = 1 (default): the LW no-scattering solver carries four g-points per lane where the shape fits
    ! (lw_solver_v7: ngpt >= 128 and a multiple of 4, nlay >= 8); 0: two per lane (lw_solver_v6). The two differ in the order of
    ! the sum over g-points only (<= 3e-7 of the flux).
    function rrnn_ctx_set_flag(ctx, name_c, val) bind(C, name="rrnn_ctx_set_flag") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: ctx
      character(kind=c_char) :: name_c(*)
      integer(c_int), value :: val
      integer(c_int) :: rc
    end function rrnn_ctx_set_flag
    ! Per-kernel device timing with CUDA events on the context's stream. rrnn_ctx_profile(ctx, 1) enables and resets;
    ! rrnn_ctx_profile_read synchronises and returns the summed duration and launch count of kernel kind 0 = NN gas optics LW, 1
    ! = LW solver, 2 = NN gas optics SW, 3 = SW solver.
    function rrnn_ctx_profile(ctx, enable) bind(C, name="rrnn_ctx_profile") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: enable
      integer(c_int) :: rc
    end function rrnn_ctx_profile
    function rrnn_ctx_profile_read(ctx, kind_i, total_ms, nlaunches) bind(C, name="rrnn_ctx_profile_read") result(rc)
      import :: c_double, c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: kind_i
      real(c_double) :: total_ms(*)
      integer(c_int) :: nlaunches(*)
      integer(c_int) :: rc
    end function rrnn_ctx_profile_read
    ! Number of kernels this context has launched since creation (bench.py's gpu_launches).
    function rrnn_ctx_launch_count(ctx) bind(C, name="rrnn_ctx_launch_count") result(rc)
      import :: c_long_long, c_ptr
      type(c_ptr), value :: ctx
      integer(c_long_long) :: rc
    end function rrnn_ctx_launch_count
    function rrnn_ctx_last_nn_kernel(ctx) bind(C, name="rrnn_ctx_last_nn_kernel") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function rrnn_ctx_last_nn_kernel
    function rrnn_ctx_nn_kernel_counts(ctx, n_tcgen05, n_ffma) bind(C, name="rrnn_ctx_nn_kernel_counts") result(rc)
      import :: c_int, c_long_long, c_ptr
      type(c_ptr), value :: ctx
      integer(c_long_long) :: n_tcgen05(*)
      integer(c_long_long) :: n_ffma(*)
      integer(c_int) :: rc
    end function rrnn_ctx_nn_kernel_counts
    ! Device memory for hosts without a CUDA binding of their own (the Fortran veneer, fortran/mo_rrnn_veneer.F90): the derived
    ! types of the reference own allocatable host arrays (tau/ssa/g, sources: rte/mo_optical_props.F90:98-192,
    ! rte/mo_source_functions.F90:26-43); here they own device buffers. Copies are ordered on the context's stream and complete
    ! before the call returns.
    function rrnn_dev_malloc(ctx, bytes, out_d) bind(C, name="rrnn_dev_malloc") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ctx
      integer(c_size_t), value :: bytes
      type(c_ptr) :: out_d
      integer(c_int) :: rc
    end function rrnn_dev_malloc
    function rrnn_dev_free(ctx, p_d) bind(C, name="rrnn_dev_free") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: p_d
      integer(c_int) :: rc
    end function rrnn_dev_free
    function rrnn_memcpy_h2d(ctx, dst_d, src, bytes) bind(C, name="rrnn_memcpy_h2d") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: dst_d
      type(c_ptr), value :: src
      integer(c_size_t), value :: bytes
      integer(c_int) :: rc
    end function rrnn_memcpy_h2d
    function rrnn_memcpy_d2h(ctx, dst, src_d, bytes) bind(C, name="rrnn_memcpy_d2h") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: dst
      type(c_ptr), value :: src_d
      integer(c_size_t), value :: bytes
      integer(c_int) :: rc
    end function rrnn_memcpy_d2h
    ! ------------------------------------------------------------------------------------------------ NN models:
    ! rrtmgp_network_type%load_netcdf, neural/mod_network_rrtmgp.F90:58-122
    function rrnn_model_load_netcdf(ctx, filename, out_h) bind(C, name="rrnn_model_load_netcdf") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: ctx
      character(kind=c_char) :: filename(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_model_load_netcdf
    ! ASCII format of network_type%load (neural/mod_network.F90:163-209) + sidecar scaling file; see INTEGRATION.md
    function rrnn_model_load_ascii(ctx, model_txt, scaling_txt, out_h) bind(C, name="rrnn_model_load_ascii") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: ctx
      character(kind=c_char) :: model_txt(*)
      character(kind=c_char) :: scaling_txt(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_model_load_ascii
    function rrnn_model_save_ascii(m, model_txt, scaling_txt) bind(C, name="rrnn_model_save_ascii") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: m
      character(kind=c_char) :: model_txt(*)
      character(kind=c_char) :: scaling_txt(*)
      integer(c_int) :: rc
    end function rrnn_model_save_ascii
    ! Build from host arrays: wpack = layer weights back to back, each row-major (n_in,n_out) (== the reference's column-major
    ! w_transposed(n_out,n_in)); input_names = nx*32 chars; ymean/ystd may be NULL.
    function rrnn_model_create(ctx, nlayers, dims, wpack, bpack, activations, xmin, xmax, ymean, ystd, input_names, &
        out_h) &
        bind(C, name="rrnn_model_create") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: nlayers
      integer(c_int) :: dims(*)
      type(c_ptr), value :: wpack
      type(c_ptr), value :: bpack
      integer(c_int) :: activations(*)
      type(c_ptr), value :: xmin
      type(c_ptr), value :: xmax
      type(c_ptr), value :: ymean
      type(c_ptr), value :: ystd
      character(kind=c_char) :: input_names(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_model_create
    function rrnn_model_destroy(m) bind(C, name="rrnn_model_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function rrnn_model_destroy
    function rrnn_model_nlayers(m) bind(C, name="rrnn_model_nlayers") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function rrnn_model_nlayers
    function rrnn_model_dims(m, dims_out) bind(C, name="rrnn_model_dims") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: dims_out(*)
      integer(c_int) :: rc
    end function rrnn_model_dims
    function rrnn_model_input_name(m, i, buf32) bind(C, name="rrnn_model_input_name") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: i
      character(kind=c_char) :: buf32(*)
      integer(c_int) :: rc
    end function rrnn_model_input_name
    function rrnn_model_activation(m, layer) bind(C, name="rrnn_model_activation") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: layer
      integer(c_int) :: rc
    end function rrnn_model_activation
    ! host copies of the parameters (for cross-checking the reader): which = 0 weights(layer), 1 bias(layer), 2 xmin, 3 xmax, 4
    ! ymean, 5 ystd. Returns the element count through n_out; data_out may be NULL.
    function rrnn_model_get(m, which, layer, data_out, n_out) bind(C, name="rrnn_model_get") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: which
      integer(c_int), value :: layer
      type(c_ptr), value :: data_out
      integer(c_int) :: n_out(*)
      integer(c_int) :: rc
    end function rrnn_model_get
    ! ------------------------------------------------------------------------------------------------ Spectral tables
    ! (ty_gas_optics_rrtmgp%load, rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326): band->g-point limits (2,nbnd) 1-based inclusive,
    ! totplnk (nPlanckTemp,nbnd) == C [nbnd][ntemp] (may be NULL for SW), solar_source (ngpt) (may be NULL for LW).
    function rrnn_kdist_create(ctx, nbnd, ngpt, band_lims_gpt, ntemp, totplnk, temp_ref_min, totplnk_delta, &
        solar_source, out_h) &
        bind(C, name="rrnn_kdist_create") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbnd
      integer(c_int), value :: ngpt
      integer(c_int) :: band_lims_gpt(*)
      integer(c_int), value :: ntemp
      type(c_ptr), value :: totplnk
      real(c_float), value :: temp_ref_min
      real(c_float), value :: totplnk_delta
      type(c_ptr), value :: solar_source
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_kdist_create
    function rrnn_kdist_destroy(kd) bind(C, name="rrnn_kdist_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: kd
      integer(c_int) :: rc
    end function rrnn_kdist_destroy
    ! ty_gas_optics_rrtmgp%set_tsi, rrtmgp/mo_gas_optics_rrtmgp.F90:1097-1120
    function rrnn_kdist_set_tsi(kd, tsi) bind(C, name="rrnn_kdist_set_tsi") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: kd
      real(c_float), value :: tsi
      integer(c_int) :: rc
    end function rrnn_kdist_set_tsi
    ! The optional solar tables of ty_gas_optics_rrtmgp%load (load_ext, rrtmgp/mo_gas_optics_rrtmgp.F90:1317-1325):
    ! solar_source_quiet / _facular / _sunspot, HOST arrays of ngpt entries.
    function rrnn_kdist_set_solar_tables(kd, solar_quiet, solar_facular, solar_sunspot) &
        bind(C, name="rrnn_kdist_set_solar_tables") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: kd
      type(c_ptr), value :: solar_quiet
      type(c_ptr), value :: solar_facular
      type(c_ptr), value :: solar_sunspot
      integer(c_int) :: rc
    end function rrnn_kdist_set_solar_tables
    ! ty_gas_optics_rrtmgp%set_solar_variability, rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1095: solar_source = quiet + (mg_index -
    ! 0.1495954) facular + (sb_index - 0.00066696) sunspot, then set_tsi(tsi) when have_tsi != 0.
    function rrnn_kdist_set_solar_variability(kd, mg_index, sb_index, have_tsi, tsi) &
        bind(C, name="rrnn_kdist_set_solar_variability") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: kd
      real(c_float), value :: mg_index
      real(c_float), value :: sb_index
      integer(c_int), value :: have_tsi
      real(c_float), value :: tsi
      integer(c_int) :: rc
    end function rrnn_kdist_set_solar_variability
    ! ty_solar_var%solar_var_ind_interp, extensions/solar_variability/mo_solar_variability.F90:91-183: facular (mg) and sunspot
    ! (sb) indices of the mean solar cycle interpolated to the cycle fraction solcycfrac in [0, 1] -- what set_solar_variability
    ! takes. avgcyc_ind is the table ty_solar_var%load keeps (:45-69), Fortran (nsolarterms = 2, nsolarfrac) == C
    ! [nsolarfrac][2], a HOST array; a host-only routine in the reference and here (no context, no device).
    function rrnn_solar_var_ind_interp(avgcyc_ind, nsolarfrac, solcycfrac, mg_index_out, sb_index_out) &
        bind(C, name="rrnn_solar_var_ind_interp") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: avgcyc_ind
      integer(c_int), value :: nsolarfrac
      real(c_float), value :: solcycfrac
      type(c_ptr), value :: mg_index_out
      type(c_ptr), value :: sb_index_out
      integer(c_int) :: rc
    end function rrnn_solar_var_ind_interp
    ! Host copy of the current solar source (ngpt).
    function rrnn_kdist_get_solar_source(kd, solar_source_out) bind(C, name="rrnn_kdist_get_solar_source") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: kd
      type(c_ptr), value :: solar_source_out
      integer(c_int) :: rc
    end function rrnn_kdist_get_solar_source
    ! optimal_angle_fit (2,nbnd) of ty_gas_optics_rrtmgp%load (:1163, 1210), a HOST array == C [nbnd][2].
    function rrnn_kdist_set_optimal_angle_fit(kd, optimal_angle_fit) bind(C, name="rrnn_kdist_set_optimal_angle_fit") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: kd
      type(c_ptr), value :: optimal_angle_fit
      integer(c_int) :: rc
    end function rrnn_kdist_set_optimal_angle_fit
    ! ty_gas_optics_rrtmgp%compute_optimal_angles, rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758: secant per column and g-point from
    ! the column transmissivity, fit(1,bnd) exp(-sum_lay tau) + fit(2,bnd). tau_d (ngpt,nlay,ncol); optimal_angles_d (ngpt,ncol)
    ! -- the layout rte_lw's lw_Ds takes here (the reference declares (ncol,ngpt), a stale upstream order).
    function rrnn_compute_optimal_angles(ctx, kd, nlay, ncol, tau_d, optimal_angles_d) &
        bind(C, name="rrnn_compute_optimal_angles") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: optimal_angles_d
      integer(c_int) :: rc
    end function rrnn_compute_optimal_angles
    ! ------------------------------------------------------------------------------------------------ Gas optics building
    ! blocks (device pointers) get_col_dry, rrtmgp/mo_gas_optics_rrtmgp.F90:1662-1707 (latitude absent)
    function rrnn_get_col_dry(ctx, ncol, nlay, vmr_h2o_d, plev_d, col_dry_d) bind(C, name="rrnn_get_col_dry") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: vmr_h2o_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: col_dry_d
      integer(c_int) :: rc
    end function rrnn_get_col_dry
    ! level-temperature interpolation, rrtmgp/mo_gas_optics_rrtmgp.F90:326-335
    function rrnn_interp_tlev(ctx, ncol, nlay, play_d, plev_d, tlay_d, tlev_d) bind(C, name="rrnn_interp_tlev") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tlev_d
      integer(c_int) :: rc
    end function rrnn_interp_tlev
    ! compute_nn_inputs, rrtmgp/mo_gas_optics_rrtmgp.F90:618-798 -> nn_inputs (ninputs,nlay,ncol)
    function rrnn_compute_nn_inputs(ctx, m, ncol, nlay, play_d, tlay_d, gases, ngas, nn_inputs_d) &
        bind(C, name="rrnn_compute_nn_inputs") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: m
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: play_d
      type(c_ptr), value :: tlay_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: nn_inputs_d
      integer(c_int) :: rc
    end function rrnn_compute_nn_inputs
    ! output_sgemm_tau / _pfrac / _lw, neural/mod_network_rrtmgp.F90:125-236, 238-317, 319-409 (x: (nx,nbatch))
    function rrnn_output_sgemm_tau(ctx, m, nbatch, x_d, coldry_d, output_d, output2_d) &
        bind(C, name="rrnn_output_sgemm_tau") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: m
      integer(c_int), value :: nbatch
      type(c_ptr), value :: x_d
      type(c_ptr), value :: coldry_d
      type(c_ptr), value :: output_d
      type(c_ptr), value :: output2_d
      integer(c_int) :: rc
    end function rrnn_output_sgemm_tau
    function rrnn_output_sgemm_pfrac(ctx, m, nbatch, x_d, output_d) bind(C, name="rrnn_output_sgemm_pfrac") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: m
      integer(c_int), value :: nbatch
      type(c_ptr), value :: x_d
      type(c_ptr), value :: output_d
      integer(c_int) :: rc
    end function rrnn_output_sgemm_pfrac
    function rrnn_output_sgemm_lw(ctx, m, nbatch, x_d, output_d) bind(C, name="rrnn_output_sgemm_lw") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: m
      integer(c_int), value :: nbatch
      type(c_ptr), value :: x_d
      type(c_ptr), value :: output_d
      integer(c_int) :: rc
    end function rrnn_output_sgemm_lw
    ! compute_Planck_source_nn, rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683: pfrac_lay_source_d holds the Planck fraction
    ! on input and lay_source on output; sfc_lay is 1-based.
    function rrnn_planck_source_nn(ctx, kd, ncol, nlay, tlay_d, tlev_d, tsfc_d, sfc_lay, sfc_source_d, &
        sfc_source_Jac_d, pfrac_lay_source_d, lev_source_d) &
        bind(C, name="rrnn_planck_source_nn") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tlev_d
      type(c_ptr), value :: tsfc_d
      integer(c_int), value :: sfc_lay
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_source_Jac_d
      type(c_ptr), value :: pfrac_lay_source_d
      type(c_ptr), value :: lev_source_d
      integer(c_int) :: rc
    end function rrnn_planck_source_nn
    ! ty_gas_optics_rrtmgp%gas_optics with neural_nets present. LW = gas_optics_int NN branch,
    ! rrtmgp/mo_gas_optics_rrtmgp.F90:239-428 (:368-411): nmodels = 2 (tau net, Planck-fraction net) or 1 ("both" net, 2*ngpt
    ! outputs); tlev_d may be NULL (-> interpolation :326-335). One fused kernel: input scaling + col_dry + MLP chain + tau /
    ! Planck-source epilogues. SW = gas_optics_ext NN branch, :433-602 (:529-573, :594-599): models[0] absorption, models[1]
    ! Rayleigh; ssa_d NULL -> 1scl request (absorption tau only); g_d NULL -> g (identically 0, :560-567) not materialised.
    function rrnn_gas_optics_lw(ctx, kd, models, nmodels, ncol, nlay, play_d, plev_d, tlay_d, tsfc_d, gases, ngas, &
        tlev_d, tau_d, lay_source_d, lev_source_d, sfc_source_d, sfc_source_Jac_d) &
        bind(C, name="rrnn_gas_optics_lw") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tsfc_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: tlev_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_source_Jac_d
      integer(c_int) :: rc
    end function rrnn_gas_optics_lw
    function rrnn_gas_optics_sw(ctx, kd, models, ncol, nlay, play_d, plev_d, tlay_d, gases, ngas, tau_d, ssa_d, g_d, &
        toa_src_d) &
        bind(C, name="rrnn_gas_optics_sw") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: toa_src_d
      integer(c_int) :: rc
    end function rrnn_gas_optics_sw
    ! The same longwave gas optics with the sources left FACTORED (no reference counterpart; this library's fused path
    ! rrnn_lw_fluxes uses it): instead of lay_source(g,l) = pfrac(g,l) B_b(T_lay(l)) and lev_source (g,l) = pfrac(g,min(l,nlay))
    ! B_b(T_lev(l)) (compute_Planck_source_nn, rrtmgp/kernels/mo_gas_optics_kernels.F90:654-672) it returns their factors:
    ! pfrac_d (ngpt,nlay,ncol) and the band Planck functions planck_lay_d (16,nlay,ncol), planck_lev_d (16,nlay+1,ncol) (rows of
    ! 16 floats, bands >= nbnd repeat the last band). 8 instead of 12 bytes per (g-point, layer) cross HBM. Two networks,
    ! tensor-core kernel only (an error otherwise); rrnn_lw_solver_noscat_compact consumes it with bit-identical fluxes.
    function rrnn_gas_optics_lw_compact(ctx, kd, models, nmodels, ncol, nlay, play_d, plev_d, tlay_d, tsfc_d, gases, &
        ngas, tlev_d, tau_d, pfrac_d, planck_lay_d, planck_lev_d, sfc_source_d, sfc_source_Jac_d) &
        bind(C, name="rrnn_gas_optics_lw_compact") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tsfc_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: tlev_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: pfrac_d
      type(c_ptr), value :: planck_lay_d
      type(c_ptr), value :: planck_lev_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_source_Jac_d
      integer(c_int) :: rc
    end function rrnn_gas_optics_lw_compact
    ! ------------------------------------------------------------------------------------------------ RTE solvers (device
    ! pointers) lw_solver_noscat_GaussQuad / lw_solver_noscat, rte/kernels/mo_rte_solver_kernels.F90:332-415, 119-330 (no
    ! rescaling, no Jacobian): Ds/weights are HOST arrays of nmus entries; inc_flux_d may be NULL (= 0).
    function rrnn_lw_solver_noscat(ctx, ngpt, nlay, ncol, top_at_1, nmus, Ds, weights, inc_flux_d, tau_d, lay_source_d, &
        lev_source_d, sfc_emis_gpt_d, sfc_source_d, flux_up_d, flux_dn_d) &
        bind(C, name="rrnn_lw_solver_noscat") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: nmus
      type(c_ptr), value :: Ds
      type(c_ptr), value :: weights
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_emis_gpt_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_solver_noscat
    ! lw_solver_noscat_GaussQuad (mo_rte_solver_kernels.F90:332-415) on the factored sources of rrnn_gas_optics_lw_compact: the
    ! solver forms lay_source / lev_source itself (one fp32 product, as mo_gas_optics_kernels.F90:654-672). kd supplies the
    ! g-point -> band map. Needs ngpt % 4 == 0, ngpt <= 512, nlay >= 8.
    function rrnn_lw_solver_noscat_compact(ctx, kd, nlay, ncol, top_at_1, nmus, Ds, weights, tau_d, pfrac_d, &
        planck_lay_d, planck_lev_d, sfc_emis_gpt_d, sfc_source_d, flux_up_d, flux_dn_d) &
        bind(C, name="rrnn_lw_solver_noscat_compact") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: nmus
      type(c_ptr), value :: Ds
      type(c_ptr), value :: weights
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: pfrac_d
      type(c_ptr), value :: planck_lay_d
      type(c_ptr), value :: planck_lev_d
      type(c_ptr), value :: sfc_emis_gpt_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_solver_noscat_compact
    ! lw_solver_noscat_GaussQuad with every option rte_lw can pass it (rte/kernels/mo_rte_solver_kernels.F90:332-415, 119-330):
    ! re-scaled scattering (do_rescaling: ssa_d, g_d, lw_transport_1rescl :1729-1795; both NULL = none), per-g-point secants
    ! lw_Ds_gpt_d (ngpt,ncol; one angle, NULL = Ds[]), g-point fluxes gpt_flux_{up,dn}_d (ngpt,nlay+1,ncol; NULL = not wanted;
    ! with one angle they hold radiances NOT multiplied by 2 pi w, as in the reference :287-291), and the surface-temperature
    ! Jacobian flux_up_Jac_d (nlay+1,ncol) from sfc_source_Jac_d (compute_Jac, rte/mo_rte_rrtmgp_config.F90:29; with one angle
    ! the sum of the un-scaled Jacobian radiances, :319). A general kernel, not the tuned benchmark path.
    function rrnn_lw_solver_noscat_ext(ctx, ngpt, nlay, ncol, top_at_1, nmus, Ds, weights, lw_Ds_gpt_d, inc_flux_d, &
        tau_d, ssa_d, g_d, lay_source_d, lev_source_d, sfc_emis_gpt_d, sfc_source_d, sfc_source_Jac_d, flux_up_d, &
        flux_dn_d, flux_up_Jac_d, gpt_flux_up_d, gpt_flux_dn_d) &
        bind(C, name="rrnn_lw_solver_noscat_ext") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: nmus
      type(c_ptr), value :: Ds
      type(c_ptr), value :: weights
      type(c_ptr), value :: lw_Ds_gpt_d
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_emis_gpt_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_source_Jac_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_up_Jac_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: gpt_flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_solver_noscat_ext
    ! lw_solver_2stream (rte/kernels/mo_rte_solver_kernels.F90:426-486: lw_two_stream :1018-1069, lw_source_2str :1112-1162,
    ! adding :1526-1637): two-stream longwave with scattering; lay_source is not used by the reference and is not an argument
    ! here; gpt_flux_{up,dn}_d (ngpt,nlay+1,ncol) optional. General kernel.
    function rrnn_lw_solver_2stream(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, tau_d, ssa_d, g_d, lev_source_d, &
        sfc_emis_gpt_d, sfc_source_d, flux_up_d, flux_dn_d, gpt_flux_up_d, gpt_flux_dn_d) &
        bind(C, name="rrnn_lw_solver_2stream") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_emis_gpt_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: gpt_flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_solver_2stream
    ! rte_lw(..., use_2stream = .true.) for ty_optical_props_2str (rte/mo_rte_lw.F90:346-361); sfc_emis_d is (nbnd,ncol).
    function rrnn_rte_lw_2stream(ctx, kd, nlay, ncol, top_at_1, inc_flux_d, tau_d, ssa_d, g_d, lev_source_d, &
        sfc_source_d, sfc_emis_d, flux_up_d, flux_dn_d, gpt_flux_up_d, gpt_flux_dn_d) &
        bind(C, name="rrnn_rte_lw_2stream") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_emis_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: gpt_flux_dn_d
      integer(c_int) :: rc
    end function rrnn_rte_lw_2stream
    ! rte_lw with its optional arguments and for ty_optical_props_2str (re-scaled solution, rte/mo_rte_lw.F90:363-384):
    ! ssa_d/g_d NULL = _1scl; lw_Ds_d (ngpt,ncol) only for _1scl and one angle (:239-249); sfc_emis_d is (nbnd,ncol).
    function rrnn_rte_lw_ext(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux_d, tau_d, ssa_d, g_d, &
        lay_source_d, lev_source_d, sfc_source_d, sfc_emis_d, lw_Ds_d, sfc_source_Jac_d, flux_up_d, flux_dn_d, &
        flux_up_Jac_d, gpt_flux_up_d, gpt_flux_dn_d) &
        bind(C, name="rrnn_rte_lw_ext") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_emis_d
      type(c_ptr), value :: lw_Ds_d
      type(c_ptr), value :: sfc_source_Jac_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_up_Jac_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: gpt_flux_dn_d
      integer(c_int) :: rc
    end function rrnn_rte_lw_ext
    ! rte_lw for ty_optical_props_1scl, rte/mo_rte_lw.F90:60-424: sfc_emis_d is (nbnd,ncol) and is expanded to g-points
    ! (:429-447); n_gauss_angles in 1..4 with the secants/weights of :113-125.
    function rrnn_rte_lw(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux_d, tau_d, lay_source_d, lev_source_d, &
        sfc_source_d, sfc_emis_d, flux_up_d, flux_dn_d) &
        bind(C, name="rrnn_rte_lw") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_emis_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_rte_lw
    ! sw_solver_2stream, rte/kernels/mo_rte_solver_kernels.F90:541-692 (two-stream :1366-1480 + adding :1526-1637).
    ! inc_flux_dif_d may be NULL (= 0, rte/mo_rte_sw.F90:191-205); g_d may be NULL (g = 0).
    function rrnn_sw_solver_2stream(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, inc_flux_dif_d, tau_d, ssa_d, g_d, &
        mu0_d, sfc_alb_dir_d, sfc_alb_dif_d, flux_up_d, flux_dn_d, flux_dir_d) &
        bind(C, name="rrnn_sw_solver_2stream") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: inc_flux_dif_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: sfc_alb_dir_d
      type(c_ptr), value :: sfc_alb_dif_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dir_d
      integer(c_int) :: rc
    end function rrnn_sw_solver_2stream
    ! sw_solver_2stream with its optional g-point fluxes (rte/kernels/mo_rte_solver_kernels.F90:541-692, save_gpt_flux):
    ! gpt_flux_{up,dn,dir}_d (ngpt,nlay+1,ncol), all three or none; gpt_flux_dn is the TOTAL downward flux (:660-663). The
    ! reference's own three sweeps (sw_two_stream_source :1366-1480, adding :1526-1637), a general kernel, not the tuned one.
    function rrnn_sw_solver_2stream_ext(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, inc_flux_dif_d, tau_d, ssa_d, g_d, &
        mu0_d, sfc_alb_dir_d, sfc_alb_dif_d, flux_up_d, flux_dn_d, flux_dir_d, gpt_flux_up_d, gpt_flux_dn_d, &
        gpt_flux_dir_d) &
        bind(C, name="rrnn_sw_solver_2stream_ext") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: inc_flux_dif_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: sfc_alb_dir_d
      type(c_ptr), value :: sfc_alb_dif_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dir_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: gpt_flux_dn_d
      type(c_ptr), value :: gpt_flux_dir_d
      integer(c_int) :: rc
    end function rrnn_sw_solver_2stream_ext
    ! rte_sw for ty_optical_props_2str, rte/mo_rte_sw.F90:48-266 (albedos per g-point, :50-61).
    function rrnn_rte_sw(ctx, ngpt, nlay, ncol, top_at_1, mu0_d, inc_flux_d, sfc_alb_dir_gpt_d, sfc_alb_dif_gpt_d, &
        inc_flux_dif_d, tau_d, ssa_d, g_d, flux_up_d, flux_dn_d, flux_dn_dir_d) &
        bind(C, name="rrnn_rte_sw") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: sfc_alb_dir_gpt_d
      type(c_ptr), value :: sfc_alb_dif_gpt_d
      type(c_ptr), value :: inc_flux_dif_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dn_dir_d
      integer(c_int) :: rc
    end function rrnn_rte_sw
    ! rte_lw / rte_sw on gas optical properties + by-band cloud optical properties (nbnd,nlay,ncol) whose
    ! `clouds%increment(atmos)` (rte/mo_optical_props.F90:714-893 -> inc_1scalar_by_1scalar_bybnd /
    ! inc_2stream_by_2stream_bybnd, rte/kernels/ mo_optical_props_kernels.F90:358-378, 453-485) has NOT been applied: the
    ! increment happens inside the solver, in registers, instead of as a read-modify-write pass over tau / ssa / g. SW: for gas
    ! properties with g == 0 (the NN gas optics); the cloud properties are taken as given (delta-scale them first if the caller
    ! does, rrnn_delta_scale_2str on the by-band arrays). Shapes the packed solvers do not take return an error that says so.
    function rrnn_rte_lw_clouds(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux_d, tau_d, lay_source_d, &
        lev_source_d, sfc_source_d, sfc_emis_d, cld_tau_bnd_d, flux_up_d, flux_dn_d) &
        bind(C, name="rrnn_rte_lw_clouds") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_emis_d
      type(c_ptr), value :: cld_tau_bnd_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_rte_lw_clouds
    function rrnn_rte_sw_clouds(ctx, kd, nlay, ncol, top_at_1, mu0_d, inc_flux_d, sfc_alb_dir_gpt_d, sfc_alb_dif_gpt_d, &
        inc_flux_dif_d, tau_d, ssa_d, cld_tau_bnd_d, cld_ssa_bnd_d, cld_g_bnd_d, flux_up_d, flux_dn_d, flux_dn_dir_d) &
        bind(C, name="rrnn_rte_sw_clouds") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: sfc_alb_dir_gpt_d
      type(c_ptr), value :: sfc_alb_dif_gpt_d
      type(c_ptr), value :: inc_flux_dif_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: cld_tau_bnd_d
      type(c_ptr), value :: cld_ssa_bnd_d
      type(c_ptr), value :: cld_g_bnd_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dn_dir_d
      integer(c_int) :: rc
    end function rrnn_rte_sw_clouds
    ! ------------------------------------------------------------------------------------------------ Cloud optics (LUT),
    ! delta-scaling, increments, heating rates ty_cloud_optics%load_lut, extensions/cloud_optics/mo_cloud_optics.F90:90-170:
    ! tables are (nsize,nbnd) == C [nbnd][nsize] for the chosen ice roughness.
    function rrnn_cloud_lut_create(ctx, nbnd, nsize_liq, nsize_ice, radliq_lwr, radliq_upr, radice_lwr, radice_upr, &
        lut_extliq, lut_ssaliq, lut_asyliq, lut_extice, lut_ssaice, lut_asyice, out_h) &
        bind(C, name="rrnn_cloud_lut_create") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbnd
      integer(c_int), value :: nsize_liq
      integer(c_int), value :: nsize_ice
      real(c_float), value :: radliq_lwr
      real(c_float), value :: radliq_upr
      real(c_float), value :: radice_lwr
      real(c_float), value :: radice_upr
      type(c_ptr), value :: lut_extliq
      type(c_ptr), value :: lut_ssaliq
      type(c_ptr), value :: lut_asyliq
      type(c_ptr), value :: lut_extice
      type(c_ptr), value :: lut_ssaice
      type(c_ptr), value :: lut_asyice
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_cloud_lut_create
    ! load_pade (extensions/cloud_optics/mo_cloud_optics.F90:178-262): Pade coefficients as stored in the coefficient files,
    ! (ncoeff, nsizereg, nbnd) with ncoeff_ext = 6 ([2/3]), ncoeff_ssa_g = 5 ([2/2]), nsizereg = 3, nbound = 4; the ice arrays
    ! for the chosen roughness. The handle is used with rrnn_cloud_optics like a LUT handle (compute_all_from_pade :650-714).
    function rrnn_cloud_pade_create(ctx, nbnd, nsizereg, ncoeff_ext, ncoeff_ssa_g, nbound, pade_extliq, pade_ssaliq, &
        pade_asyliq, pade_extice, pade_ssaice, pade_asyice, sizreg_extliq, sizreg_ssaliq, sizreg_asyliq, sizreg_extice, &
        sizreg_ssaice, sizreg_asyice, out_h) &
        bind(C, name="rrnn_cloud_pade_create") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbnd
      integer(c_int), value :: nsizereg
      integer(c_int), value :: ncoeff_ext
      integer(c_int), value :: ncoeff_ssa_g
      integer(c_int), value :: nbound
      type(c_ptr), value :: pade_extliq
      type(c_ptr), value :: pade_ssaliq
      type(c_ptr), value :: pade_asyliq
      type(c_ptr), value :: pade_extice
      type(c_ptr), value :: pade_ssaice
      type(c_ptr), value :: pade_asyice
      type(c_ptr), value :: sizreg_extliq
      type(c_ptr), value :: sizreg_ssaliq
      type(c_ptr), value :: sizreg_asyliq
      type(c_ptr), value :: sizreg_extice
      type(c_ptr), value :: sizreg_ssaice
      type(c_ptr), value :: sizreg_asyice
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_cloud_pade_create
    function rrnn_cloud_lut_destroy(lut) bind(C, name="rrnn_cloud_lut_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: lut
      integer(c_int) :: rc
    end function rrnn_cloud_lut_destroy
    ! ty_cloud_optics%cloud_optics, mo_cloud_optics.F90:354-535 (compute_all_from_table :603-645): by-band (nbnd,nlay,ncol)
    ! outputs; ssa_d = g_d = NULL -> 1scl absorption optical depth (:505-513).
    function rrnn_cloud_optics(ctx, lut, ncol, nlay, clwp_d, ciwp_d, reliq_d, reice_d, tau_d, ssa_d, g_d) &
        bind(C, name="rrnn_cloud_optics") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: lut
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: clwp_d
      type(c_ptr), value :: ciwp_d
      type(c_ptr), value :: reliq_d
      type(c_ptr), value :: reice_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      integer(c_int) :: rc
    end function rrnn_cloud_optics
    ! McICA sampling (extensions/cloud_optics/mo_cloud_sampling.F90): sampled_mask_max_ran :107-170 (overlap_param_d NULL) and
    ! sampled_mask_exp_ran :176-286; randoms_d and the mask (1 byte per element) are (ngpt,nlay,ncol), cloud_frac_d (nlay,ncol),
    ! overlap_param_d (nlay-1,ncol) -- this fork's layout throughout (the module itself still declares (ncol,nlay,ngpt)).
    function rrnn_sampled_mask(ctx, ngpt, nlay, ncol, randoms_d, cloud_frac_d, overlap_param_d, cloud_mask_d) &
        bind(C, name="rrnn_sampled_mask") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ngpt
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      type(c_ptr), value :: randoms_d
      type(c_ptr), value :: cloud_frac_d
      type(c_ptr), value :: overlap_param_d
      type(c_ptr), value :: cloud_mask_d
      integer(c_int) :: rc
    end function rrnn_sampled_mask
    ! draw_samples / apply_cloud_mask (:38-101, 292-308): by-band cloud properties (nbnd,nlay,ncol) -> sampled by g-point
    ! (ngpt,nlay,ncol), zero where the mask is false; ssa/g NULL for ty_optical_props_1scl.
    function rrnn_draw_samples(ctx, kd, nlay, ncol, cloud_mask_d, tau_bnd_d, ssa_bnd_d, g_bnd_d, tau_gpt_d, ssa_gpt_d, &
        g_gpt_d) &
        bind(C, name="rrnn_draw_samples") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      type(c_ptr), value :: cloud_mask_d
      type(c_ptr), value :: tau_bnd_d
      type(c_ptr), value :: ssa_bnd_d
      type(c_ptr), value :: g_bnd_d
      type(c_ptr), value :: tau_gpt_d
      type(c_ptr), value :: ssa_gpt_d
      type(c_ptr), value :: g_gpt_d
      integer(c_int) :: rc
    end function rrnn_draw_samples
    ! delta_scale_2str_k, rte/kernels/mo_optical_props_kernels.F90:72-93 (n = number of elements)
    function rrnn_delta_scale_2str(ctx, n, tau_d, ssa_d, g_d) bind(C, name="rrnn_delta_scale_2str") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ctx
      integer(c_size_t), value :: n
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      integer(c_int) :: rc
    end function rrnn_delta_scale_2str
    ! inc_1scalar_by_1scalar_bybnd :358-378 and inc_2stream_by_2stream_bybnd :453-485; gpt_lims from kd. /
    function rrnn_increment_1scl_bybnd(ctx, kd, nlay, ncol, tau1_d, tau2_d) bind(C, name="rrnn_increment_1scl_bybnd") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      type(c_ptr), value :: tau1_d
      type(c_ptr), value :: tau2_d
      integer(c_int) :: rc
    end function rrnn_increment_1scl_bybnd
    function rrnn_increment_2str_bybnd(ctx, kd, nlay, ncol, tau1_d, ssa1_d, g1_d, tau2_d, ssa2_d, g2_d) &
        bind(C, name="rrnn_increment_2str_bybnd") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      type(c_ptr), value :: tau1_d
      type(c_ptr), value :: ssa1_d
      type(c_ptr), value :: g1_d
      type(c_ptr), value :: tau2_d
      type(c_ptr), value :: ssa2_d
      type(c_ptr), value :: g2_d
      integer(c_int) :: rc
    end function rrnn_increment_2str_bybnd
    ! compute_heating_rate, extensions/mo_heating_rates.F90:26-54 [K/s] (this fork's (nlay+1,ncol) layout) and
    ! calc_heating_rate, examples/rrtmgp-nn-training/rrtmgp_lw_eval_nn_rfmip.F90:624-653 [K/day].
    function rrnn_heating_rate(ctx, ncol, nlay, flux_up_d, flux_dn_d, plev_d, heating_rate_d) &
        bind(C, name="rrnn_heating_rate") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: heating_rate_d
      integer(c_int) :: rc
    end function rrnn_heating_rate
    function rrnn_calc_heating_rate(ctx, ncol, nlay, flux_up_d, flux_dn_d, plev_d, hr_K_day_d) &
        bind(C, name="rrnn_calc_heating_rate") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: hr_K_day_d
      integer(c_int) :: rc
    end function rrnn_calc_heating_rate
    ! ty_fluxes_byband%reduce, extensions/mo_fluxes_byband.F90:41-131. sum_byband (mo_fluxes_byband_kernels.F90:33-51): g-point
    ! fluxes (ngpt,nlev,ncol) -> by-band (nbnd,nlev,ncol), summed in g-point order; net_byband_full (:56-78): by-band sum of
    ! (down - up). Band limits from kd. This fork's layout (g-point / band fastest).
    function rrnn_sum_byband(ctx, kd, nlev, ncol, gpt_flux_d, bnd_flux_d) bind(C, name="rrnn_sum_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlev
      integer(c_int), value :: ncol
      type(c_ptr), value :: gpt_flux_d
      type(c_ptr), value :: bnd_flux_d
      integer(c_int) :: rc
    end function rrnn_sum_byband
    function rrnn_net_byband(ctx, kd, nlev, ncol, gpt_flux_dn_d, gpt_flux_up_d, bnd_flux_net_d) &
        bind(C, name="rrnn_net_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlev
      integer(c_int), value :: ncol
      type(c_ptr), value :: gpt_flux_dn_d
      type(c_ptr), value :: gpt_flux_up_d
      type(c_ptr), value :: bnd_flux_net_d
      integer(c_int) :: rc
    end function rrnn_net_byband
    ! rte_lw / rte_sw with ty_fluxes_byband (extensions/mo_fluxes_byband.F90:41-131) WITHOUT g-point fluxes: broadband and
    ! by-band fluxes (nbnd,nlay+1,ncol) both come out of the tuned solver, whose per-level sums pass through the band sums
    ! anyway. Needs bands of 16 aligned g-points (every RRTMGP k-distribution) and a shape the packed solver takes; otherwise an
    ! error that says so, and the general path (g-point fluxes + rrnn_sum_byband) is the one to use. Summation order differs
    ! from sum_byband's (in g-point order), so these agree with it to rounding, not bit for bit. LW with ONE quadrature angle:
    ! the by-band values are sums of the reference's un-scaled g-point radiances, as there (quirk Q3,
    ! rte/kernels/mo_rte_solver_kernels.F90:284-291); the broadband fluxes are fluxes.
    function rrnn_rte_lw_byband(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, inc_flux_d, tau_d, lay_source_d, &
        lev_source_d, sfc_source_d, sfc_emis_d, flux_up_d, flux_dn_d, bnd_flux_up_d, bnd_flux_dn_d) &
        bind(C, name="rrnn_rte_lw_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: lay_source_d
      type(c_ptr), value :: lev_source_d
      type(c_ptr), value :: sfc_source_d
      type(c_ptr), value :: sfc_emis_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: bnd_flux_up_d
      type(c_ptr), value :: bnd_flux_dn_d
      integer(c_int) :: rc
    end function rrnn_rte_lw_byband
    function rrnn_rte_sw_byband(ctx, kd, nlay, ncol, top_at_1, mu0_d, inc_flux_d, sfc_alb_dir_gpt_d, sfc_alb_dif_gpt_d, &
        inc_flux_dif_d, tau_d, ssa_d, g_d, flux_up_d, flux_dn_d, flux_dn_dir_d, bnd_flux_up_d, bnd_flux_dn_d, &
        bnd_flux_dn_dir_d) &
        bind(C, name="rrnn_rte_sw_byband") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      integer(c_int), value :: nlay
      integer(c_int), value :: ncol
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: inc_flux_d
      type(c_ptr), value :: sfc_alb_dir_gpt_d
      type(c_ptr), value :: sfc_alb_dif_gpt_d
      type(c_ptr), value :: inc_flux_dif_d
      type(c_ptr), value :: tau_d
      type(c_ptr), value :: ssa_d
      type(c_ptr), value :: g_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dn_dir_d
      type(c_ptr), value :: bnd_flux_up_d
      type(c_ptr), value :: bnd_flux_dn_d
      type(c_ptr), value :: bnd_flux_dn_dir_d
      integer(c_int) :: rc
    end function rrnn_rte_sw_byband
    ! net = down - up over n elements: net_byband_precalc (mo_fluxes_byband_kernels.F90:80-86) and the broadband flux_net of
    ! ty_fluxes_broadband%reduce (rte/mo_fluxes.F90, net_broadband_precalc).
    function rrnn_net_flux(ctx, n, flux_dn_d, flux_up_d, flux_net_d) bind(C, name="rrnn_net_flux") result(rc)
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ctx
      integer(c_size_t), value :: n
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_net_d
      integer(c_int) :: rc
    end function rrnn_net_flux
    ! ------------------------------------------------------------------------------------------------ Whole-path drivers with
    ! HOST buffers: what one iteration of the reference drivers' block loop does
    ! (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446, rrtmgp_rfmip_sw.F90:356-465): gas_optics -> rte_lw / rte_sw, here
    ! for all columns at once, in column chunks that fit the device workspace, H2D/D2H overlapped with compute. sfc_emis/sfc_alb
    ! are per column (spectrally constant, as in the RFMIP drivers); mu0 <= 0 marks night columns (fluxes zeroed,
    ! rrtmgp_rfmip_sw.F90:458-463). tsi_scale may be NULL; else the per-column TSI renormalisation of :409-416 is applied. Gas
    ! conc pointers are HOST pointers.
    function rrnn_lw_fluxes_host(ctx, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play, plev, tlay, &
        tlev, tsfc, sfc_emis, gases, ngas, flux_up, flux_dn) &
        bind(C, name="rrnn_lw_fluxes_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: tlev
      type(c_ptr), value :: tsfc
      type(c_ptr), value :: sfc_emis
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      integer(c_int) :: rc
    end function rrnn_lw_fluxes_host
    function rrnn_sw_fluxes_host(ctx, kd, models, ncol, nlay, top_at_1, play, plev, tlay, mu0, sfc_alb, tsi, gases, &
        ngas, flux_up, flux_dn, flux_dn_dir) &
        bind(C, name="rrnn_sw_fluxes_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: mu0
      type(c_ptr), value :: sfc_alb
      type(c_ptr), value :: tsi
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      type(c_ptr), value :: flux_dn_dir
      integer(c_int) :: rc
    end function rrnn_sw_fluxes_host
    ! Same path with DEVICE buffers (inputs already resident; used for the kernel-only throughput number).
    function rrnn_lw_fluxes(ctx, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play_d, plev_d, tlay_d, &
        tlev_d, tsfc_d, sfc_emis_d, gases, ngas, flux_up_d, flux_dn_d) &
        bind(C, name="rrnn_lw_fluxes") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tlev_d
      type(c_ptr), value :: tsfc_d
      type(c_ptr), value :: sfc_emis_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_fluxes
    function rrnn_sw_fluxes(ctx, kd, models, ncol, nlay, top_at_1, play_d, plev_d, tlay_d, mu0_d, sfc_alb_d, tsi_d, &
        gases, ngas, flux_up_d, flux_dn_d, flux_dn_dir_d) &
        bind(C, name="rrnn_sw_fluxes") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: sfc_alb_d
      type(c_ptr), value :: tsi_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dn_dir_d
      integer(c_int) :: rc
    end function rrnn_sw_fluxes
    ! All-sky whole-path drivers: one iteration of examples/all-sky/rrtmgp_allsky.F90:366-446 -- cloud_optics (LUT or Pade
    ! handle, by band), gas_optics(neural_nets=), [delta_scale,] increment, rte_lw / rte_sw -- for all columns of the call. clwp
    ! / ciwp / reliq / reice are (nlay,ncol); the other arguments are those of the clear-sky drivers above. The cloud increment
    ! is NOT a pass over the (ngpt,nlay,ncol) arrays: the by-band cloud properties are added to the gas properties inside the
    ! solvers.
    function rrnn_lw_fluxes_allsky(ctx, kd, models, nmodels, cloud_optics, ncol, nlay, top_at_1, n_gauss_angles, &
        play_d, plev_d, tlay_d, tlev_d, tsfc_d, sfc_emis_d, gases, ngas, clwp_d, ciwp_d, reliq_d, reice_d, flux_up_d, &
        flux_dn_d) &
        bind(C, name="rrnn_lw_fluxes_allsky") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      type(c_ptr), value :: cloud_optics
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: tlev_d
      type(c_ptr), value :: tsfc_d
      type(c_ptr), value :: sfc_emis_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: clwp_d
      type(c_ptr), value :: ciwp_d
      type(c_ptr), value :: reliq_d
      type(c_ptr), value :: reice_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      integer(c_int) :: rc
    end function rrnn_lw_fluxes_allsky
    function rrnn_sw_fluxes_allsky(ctx, kd, models, cloud_optics, ncol, nlay, top_at_1, play_d, plev_d, tlay_d, mu0_d, &
        sfc_alb_d, tsi_d, gases, ngas, clwp_d, ciwp_d, reliq_d, reice_d, flux_up_d, flux_dn_d, flux_dn_dir_d) &
        bind(C, name="rrnn_sw_fluxes_allsky") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      type(c_ptr), value :: cloud_optics
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: play_d
      type(c_ptr), value :: plev_d
      type(c_ptr), value :: tlay_d
      type(c_ptr), value :: mu0_d
      type(c_ptr), value :: sfc_alb_d
      type(c_ptr), value :: tsi_d
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: clwp_d
      type(c_ptr), value :: ciwp_d
      type(c_ptr), value :: reliq_d
      type(c_ptr), value :: reice_d
      type(c_ptr), value :: flux_up_d
      type(c_ptr), value :: flux_dn_d
      type(c_ptr), value :: flux_dn_dir_d
      integer(c_int) :: rc
    end function rrnn_sw_fluxes_allsky
    function rrnn_lw_fluxes_allsky_host(ctx, kd, models, nmodels, cloud_optics, ncol, nlay, top_at_1, n_gauss_angles, &
        play, plev, tlay, tlev, tsfc, sfc_emis, gases, ngas, clwp, ciwp, reliq, reice, flux_up, flux_dn) &
        bind(C, name="rrnn_lw_fluxes_allsky_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      integer(c_int), value :: nmodels
      type(c_ptr), value :: cloud_optics
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: tlev
      type(c_ptr), value :: tsfc
      type(c_ptr), value :: sfc_emis
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: clwp
      type(c_ptr), value :: ciwp
      type(c_ptr), value :: reliq
      type(c_ptr), value :: reice
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      integer(c_int) :: rc
    end function rrnn_lw_fluxes_allsky_host
    function rrnn_sw_fluxes_allsky_host(ctx, kd, models, cloud_optics, ncol, nlay, top_at_1, play, plev, tlay, mu0, &
        sfc_alb, tsi, gases, ngas, clwp, ciwp, reliq, reice, flux_up, flux_dn, flux_dn_dir) &
        bind(C, name="rrnn_sw_fluxes_allsky_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: ctx
      type(c_ptr), value :: kd
      type(c_ptr) :: models(*)
      type(c_ptr), value :: cloud_optics
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: mu0
      type(c_ptr), value :: sfc_alb
      type(c_ptr), value :: tsi
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: clwp
      type(c_ptr), value :: ciwp
      type(c_ptr), value :: reliq
      type(c_ptr), value :: reice
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      type(c_ptr), value :: flux_dn_dir
      integer(c_int) :: rc
    end function rrnn_sw_fluxes_allsky_host
    ! Column chunk used by the drivers above (0 = automatic from free device memory).
    function rrnn_ctx_set_chunk_columns(ctx, ncol_chunk) bind(C, name="rrnn_ctx_set_chunk_columns") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol_chunk
      integer(c_int) :: rc
    end function rrnn_ctx_set_chunk_columns
    function rrnn_nc_open(path, out_h) bind(C, name="rrnn_nc_open") result(rc)
      import :: c_char, c_int, c_ptr
      character(kind=c_char) :: path(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_nc_open
    function rrnn_nc_create(path, out_h) bind(C, name="rrnn_nc_create") result(rc)
      import :: c_char, c_int, c_ptr
      character(kind=c_char) :: path(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_nc_create
    function rrnn_nc_close(f) bind(C, name="rrnn_nc_close") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: f
      integer(c_int) :: rc
    end function rrnn_nc_close
    function rrnn_nc_var_exists(f, name_c) bind(C, name="rrnn_nc_var_exists") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: f
      character(kind=c_char) :: name_c(*)
      integer(c_int) :: rc
    end function rrnn_nc_var_exists
    ! 1 / 0
    function rrnn_nc_inq_var(f, name_c, ndims, shape) bind(C, name="rrnn_nc_inq_var") result(rc)
      import :: c_char, c_int, c_long_long, c_ptr
      type(c_ptr), value :: f
      character(kind=c_char) :: name_c(*)
      integer(c_int) :: ndims(*)
      integer(c_long_long) :: shape(*)
      integer(c_int) :: rc
    end function rrnn_nc_inq_var
    ! any numeric variable, converted to float (real(wp) read_field)
    function rrnn_nc_get_var_float(f, name_c, data_out, n) bind(C, name="rrnn_nc_get_var_float") result(rc)
      import :: c_char, c_int, c_ptr, c_size_t
      type(c_ptr), value :: f
      character(kind=c_char) :: name_c(*)
      type(c_ptr), value :: data_out
      integer(c_size_t), value :: n
      integer(c_int) :: rc
    end function rrnn_nc_get_var_float
    function rrnn_nc_get_att_text(f, var, att, buf, nbuf) bind(C, name="rrnn_nc_get_att_text") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: f
      character(kind=c_char) :: var(*)
      character(kind=c_char) :: att(*)
      character(kind=c_char) :: buf(*)
      integer(c_int), value :: nbuf
      integer(c_int) :: rc
    end function rrnn_nc_get_att_text
    function rrnn_nc_def_dim(f, name_c, len_i, dimid) bind(C, name="rrnn_nc_def_dim") result(rc)
      import :: c_char, c_int, c_long_long, c_ptr
      type(c_ptr), value :: f
      character(kind=c_char) :: name_c(*)
      integer(c_long_long), value :: len_i
      integer(c_int) :: dimid(*)
      integer(c_int) :: rc
    end function rrnn_nc_def_dim
    ! create_var + write_field: a float variable over already-defined dimensions; units may be NULL
    function rrnn_nc_put_var_float(f, name_c, ndims, dimids, data_p, units) bind(C, name="rrnn_nc_put_var_float") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: f
      character(kind=c_char) :: name_c(*)
      integer(c_int), value :: ndims
      integer(c_int) :: dimids(*)
      type(c_ptr), value :: data_p
      character(kind=c_char) :: units(*)
      integer(c_int) :: rc
    end function rrnn_nc_put_var_float
    function rrnn_multi_create(ndev, devices, out_h) bind(C, name="rrnn_multi_create") result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: ndev
      integer(c_int) :: devices(*)
      type(c_ptr) :: out_h
      integer(c_int) :: rc
    end function rrnn_multi_create
    function rrnn_multi_destroy(m) bind(C, name="rrnn_multi_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function rrnn_multi_destroy
    function rrnn_multi_ndev(m) bind(C, name="rrnn_multi_ndev") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function rrnn_multi_ndev
    function rrnn_multi_ctx(m, i) bind(C, name="rrnn_multi_ctx") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: i
      type(c_ptr) :: rc
    end function rrnn_multi_ctx
    ! the i-th device's context (flags, profiling)
    function rrnn_multi_set_flag(m, name_c, val) bind(C, name="rrnn_multi_set_flag") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: m
      character(kind=c_char) :: name_c(*)
      integer(c_int), value :: val
      integer(c_int) :: rc
    end function rrnn_multi_set_flag
    ! rrnn_ctx_set_flag on every device rrtmgp_network_type%load_netcdf (neural/mod_network_rrtmgp.F90:58-122) onto every
    ! device; returns an id for the calls below
    function rrnn_multi_model_load_netcdf(m, filename, model_id) bind(C, name="rrnn_multi_model_load_netcdf") result(rc)
      import :: c_char, c_int, c_ptr
      type(c_ptr), value :: m
      character(kind=c_char) :: filename(*)
      integer(c_int) :: model_id(*)
      integer(c_int) :: rc
    end function rrnn_multi_model_load_netcdf
    ! rrnn_kdist_create on every device; returns an id
    function rrnn_multi_kdist_create(m, nbnd, ngpt, band_lims_gpt, ntemp, totplnk, temp_ref_min, totplnk_delta, &
        solar_source, kdist_id) &
        bind(C, name="rrnn_multi_kdist_create") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: nbnd
      integer(c_int), value :: ngpt
      integer(c_int) :: band_lims_gpt(*)
      integer(c_int), value :: ntemp
      type(c_ptr), value :: totplnk
      real(c_float), value :: temp_ref_min
      real(c_float), value :: totplnk_delta
      type(c_ptr), value :: solar_source
      integer(c_int) :: kdist_id(*)
      integer(c_int) :: rc
    end function rrnn_multi_kdist_create
    function rrnn_multi_kdist_set_tsi(m, kdist_id, tsi) bind(C, name="rrnn_multi_kdist_set_tsi") result(rc)
      import :: c_float, c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int), value :: kdist_id
      real(c_float), value :: tsi
      integer(c_int) :: rc
    end function rrnn_multi_kdist_set_tsi
    ! rrnn_lw_fluxes_host / rrnn_sw_fluxes_host over all devices (same arguments, ids instead of handles)
    function rrnn_multi_lw_fluxes_host(m, kdist_id, model_ids, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play, &
        plev, tlay, tlev, tsfc, sfc_emis, gases, ngas, flux_up, flux_dn) &
        bind(C, name="rrnn_multi_lw_fluxes_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: m
      integer(c_int), value :: kdist_id
      integer(c_int) :: model_ids(*)
      integer(c_int), value :: nmodels
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      integer(c_int), value :: n_gauss_angles
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: tlev
      type(c_ptr), value :: tsfc
      type(c_ptr), value :: sfc_emis
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      integer(c_int) :: rc
    end function rrnn_multi_lw_fluxes_host
    function rrnn_multi_sw_fluxes_host(m, kdist_id, model_ids, ncol, nlay, top_at_1, play, plev, tlay, mu0, sfc_alb, &
        tsi, gases, ngas, flux_up, flux_dn, flux_dn_dir) &
        bind(C, name="rrnn_multi_sw_fluxes_host") result(rc)
      import :: c_int, c_ptr, rrnn_gas_t
      type(c_ptr), value :: m
      integer(c_int), value :: kdist_id
      integer(c_int) :: model_ids(*)
      integer(c_int), value :: ncol
      integer(c_int), value :: nlay
      integer(c_int), value :: top_at_1
      type(c_ptr), value :: play
      type(c_ptr), value :: plev
      type(c_ptr), value :: tlay
      type(c_ptr), value :: mu0
      type(c_ptr), value :: sfc_alb
      type(c_ptr), value :: tsi
      type(rrnn_gas_t) :: gases(*)
      integer(c_int), value :: ngas
      type(c_ptr), value :: flux_up
      type(c_ptr), value :: flux_dn
      type(c_ptr), value :: flux_dn_dir
      integer(c_int) :: rc
    end function rrnn_multi_sw_fluxes_host
  end interface

contains

  ! the reference's error convention: character(len=128), empty = success (rte/mo_rte_lw.F90:88, 140)
  function rrnn_error_msg(rc) result(error_msg)
    integer(c_int), intent(in) :: rc
    character(len=128)         :: error_msg
    type(c_ptr) :: p
    character(kind=c_char), pointer :: s(:)
    integer :: i
    error_msg = ""
    if (rc == 0) return
    p = rrnn_last_error()
    if (.not. c_associated(p)) then
      error_msg = "librrnn_b200: error"
      return
    end if
    call c_f_pointer(p, s, [128])
    do i = 1, 128
      if (s(i) == c_null_char) exit
      error_msg(i:i) = s(i)
    end do
  end function rrnn_error_msg

end module mo_rrnn_c_binding
