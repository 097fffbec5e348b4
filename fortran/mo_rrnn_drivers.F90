! Fused whole-path calls for a Fortran host: the body of one block-loop iteration of examples/rfmip-clear-sky/
! rrtmgp_rfmip_{lw,sw}.F90 (gas_optics with neural_nets -> rte_lw / rte_sw) as ONE call on host arrays, for all columns at
! once, on one device (rrnn_lw / rrnn_sw) or on every device of the node (rrnn_lw_multi / rrnn_sw_multi -- the reference's
! OpenMP-over-blocks loop, rrtmgp_rfmip_lw.F90:364-368, with GPUs in the place of threads).  The intermediate optical
! properties never leave the device.  The type-level veneer with the reference's own names is mo_rrnn_veneer.F90; these
! calls are the fast path next to it.  Argument meaning and the character(len=128) error convention are the reference's.
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (see mo_rrnn_c_binding.F90); checked by tests/test_fortran_cpu.py.
module mo_rrnn_drivers
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, c_str
  use mo_gas_optics_rrtmgp, only: ty_gas_optics_rrtmgp
  use mod_network_rrtmgp, only: rrtmgp_network_type
  use mo_gas_concentrations, only: ty_gas_concs
  use mo_cloud_optics, only: ty_cloud_optics
  implicit none
  private
  public :: rrnn_lw, rrnn_sw, rrnn_lw_allsky, rrnn_sw_allsky, ty_rrnn_multi

  ! One process, N devices (rrnn_multi_*, include/rrnn.h)
  type :: ty_rrnn_multi
    type(c_ptr) :: h = c_null_ptr
  contains
    procedure :: init        => multi_init            ! devices(:) optional: default every visible device
    procedure :: load_netcdf => multi_load_netcdf     ! -> model id
    procedure :: load_lw     => multi_load_lw         ! -> kdist id
    procedure :: load_sw     => multi_load_sw
    procedure :: lw          => multi_lw
    procedure :: sw          => multi_sw
    procedure :: finalize    => multi_finalize
  end type ty_rrnn_multi

contains

  ! host-side rrnn_gas_t array: conc pointers are HOST addresses for the *_host entry points
  subroutine host_gases(gas_desc, gases)
    type(ty_gas_concs), target,    intent(in)  :: gas_desc
    type(rrnn_gas_t), allocatable, intent(out) :: gases(:)
    integer :: i, k
    allocate(gases(gas_desc%get_num_gases()))
    do i = 1, size(gases)
      gases(i)%name = c_null_char
      do k = 1, len_trim(gas_desc%gas_name(i))
        gases(i)%name(k) = gas_desc%gas_name(i)(k:k)
      end do
      gases(i)%conc = c_null_ptr
      gases(i)%value = 0._wp
      gases(i)%ndims = 0
      if (.not. allocated(gas_desc%concs(i)%conc)) cycle
      if (size(gas_desc%concs(i)%conc) == 1) then
        gases(i)%value = gas_desc%concs(i)%conc(1, 1)
      else
        gases(i)%conc = c_loc(gas_desc%concs(i)%conc)
        gases(i)%ndims = merge(1, 2, size(gas_desc%concs(i)%conc, 2) == 1 .and. gas_desc%ncol /= 1)
      end if
    end do
  end subroutine host_gases

  function rrnn_lw(k_dist, neural_nets, play, plev, tlay, tsfc, sfc_emis, gas_desc, top_at_1, flux_up, flux_dn, &
                   tlev, n_gauss_angles) result(error_msg)
    type(ty_gas_optics_rrtmgp), intent(in) :: k_dist
    type(rrtmgp_network_type),  intent(in) :: neural_nets(:)
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), tsfc(:), sfc_emis(:)    ! sfc_emis per column
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:)                               ! (nlay+1, ncol)
    real(wp), contiguous, target, optional, intent(in) :: tlev(:,:)
    integer,                      optional, intent(in) :: n_gauss_angles
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    type(c_ptr) :: models(2), p_tlev
    integer :: i, nang
    models = c_null_ptr
    do i = 1, min(size(neural_nets), 2)
      models(i) = neural_nets(i)%handle
    end do
    p_tlev = c_null_ptr
    if (present(tlev)) p_tlev = c_loc(tlev)
    nang = 1
    if (present(n_gauss_angles)) nang = n_gauss_angles
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_lw_fluxes_host(rrnn_ctx(), k_dist%kd, models, int(size(neural_nets), c_int), &
                  int(size(play, 2), c_int), int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), int(nang, c_int), &
                  c_loc(play), c_loc(plev), c_loc(tlay), p_tlev, c_loc(tsfc), c_loc(sfc_emis), gases, int(size(gases), c_int), &
                  c_loc(flux_up), c_loc(flux_dn)))
  end function rrnn_lw

  function rrnn_sw(k_dist, neural_nets, play, plev, tlay, mu0, sfc_alb, gas_desc, top_at_1, flux_up, flux_dn, flux_dn_dir, &
                   tsi_scale) result(error_msg)
    type(ty_gas_optics_rrtmgp), intent(in) :: k_dist
    type(rrtmgp_network_type),  intent(in) :: neural_nets(2)
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), mu0(:), sfc_alb(:)
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:), flux_dn_dir(:,:)
    real(wp), contiguous, target, optional, intent(in) :: tsi_scale(:)        ! rrtmgp_rfmip_sw.F90:409-416
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    type(c_ptr) :: models(2), p_tsi
    models(1) = neural_nets(1)%handle
    models(2) = neural_nets(2)%handle
    p_tsi = c_null_ptr
    if (present(tsi_scale)) p_tsi = c_loc(tsi_scale)
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_sw_fluxes_host(rrnn_ctx(), k_dist%kd, models, int(size(play, 2), c_int), &
                  int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), c_loc(play), c_loc(plev), c_loc(tlay), &
                  c_loc(mu0), c_loc(sfc_alb), p_tsi, gases, int(size(gases), c_int), c_loc(flux_up), c_loc(flux_dn), &
                  c_loc(flux_dn_dir)))
  end function rrnn_sw

  ! ---------------------------------------------------------------------------------------------------- all-sky
  ! One iteration of examples/all-sky/rrtmgp_allsky.F90:366-446 for all columns: cloud_optics (by band), gas_optics(neural_nets=),
  ! [delta_scale,] increment, rte -- the increment happens inside the solvers (no pass over the (ngpt,nlay,ncol) arrays).
  function rrnn_lw_allsky(k_dist, neural_nets, cloud_optics, play, plev, tlay, tsfc, sfc_emis, gas_desc, clwp, ciwp, reliq, reice, &
                          top_at_1, flux_up, flux_dn, tlev, n_gauss_angles) result(error_msg)
    type(ty_gas_optics_rrtmgp), intent(in) :: k_dist
    type(rrtmgp_network_type),  intent(in) :: neural_nets(:)
    type(ty_cloud_optics),      intent(in) :: cloud_optics
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), tsfc(:), sfc_emis(:)
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    real(wp), contiguous, target, intent(in)  :: clwp(:,:), ciwp(:,:), reliq(:,:), reice(:,:)            ! (nlay, ncol)
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:)
    real(wp), contiguous, target, optional, intent(in) :: tlev(:,:)
    integer,                      optional, intent(in) :: n_gauss_angles
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    type(c_ptr) :: models(2), p_tlev
    integer :: i, nang
    models = c_null_ptr
    do i = 1, min(size(neural_nets), 2)
      models(i) = neural_nets(i)%handle
    end do
    p_tlev = c_null_ptr
    if (present(tlev)) p_tlev = c_loc(tlev)
    nang = 1
    if (present(n_gauss_angles)) nang = n_gauss_angles
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_lw_fluxes_allsky_host(rrnn_ctx(), k_dist%kd, models, int(size(neural_nets), c_int), &
                  cloud_optics%lut, int(size(play, 2), c_int), int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), &
                  int(nang, c_int), c_loc(play), c_loc(plev), c_loc(tlay), p_tlev, c_loc(tsfc), c_loc(sfc_emis), gases, &
                  int(size(gases), c_int), c_loc(clwp), c_loc(ciwp), c_loc(reliq), c_loc(reice), c_loc(flux_up), c_loc(flux_dn)))
  end function rrnn_lw_allsky

  function rrnn_sw_allsky(k_dist, neural_nets, cloud_optics, play, plev, tlay, mu0, sfc_alb, gas_desc, clwp, ciwp, reliq, reice, &
                          top_at_1, flux_up, flux_dn, flux_dn_dir, tsi_scale) result(error_msg)
    type(ty_gas_optics_rrtmgp), intent(in) :: k_dist
    type(rrtmgp_network_type),  intent(in) :: neural_nets(2)
    type(ty_cloud_optics),      intent(in) :: cloud_optics
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), mu0(:), sfc_alb(:)
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    real(wp), contiguous, target, intent(in)  :: clwp(:,:), ciwp(:,:), reliq(:,:), reice(:,:)
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:), flux_dn_dir(:,:)
    real(wp), contiguous, target, optional, intent(in) :: tsi_scale(:)
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    type(c_ptr) :: models(2), p_tsi
    models(1) = neural_nets(1)%handle
    models(2) = neural_nets(2)%handle
    p_tsi = c_null_ptr
    if (present(tsi_scale)) p_tsi = c_loc(tsi_scale)
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_sw_fluxes_allsky_host(rrnn_ctx(), k_dist%kd, models, cloud_optics%lut, int(size(play, 2), c_int), &
                  int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), c_loc(play), c_loc(plev), c_loc(tlay), c_loc(mu0), &
                  c_loc(sfc_alb), p_tsi, gases, int(size(gases), c_int), c_loc(clwp), c_loc(ciwp), c_loc(reliq), c_loc(reice), &
                  c_loc(flux_up), c_loc(flux_dn), c_loc(flux_dn_dir)))
  end function rrnn_sw_allsky

  ! ---------------------------------------------------------------------------------------------------- N devices
  function multi_init(this, devices) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    integer, optional,    intent(in)    :: devices(:)
    character(len=128) :: error_msg
    integer(c_int), allocatable :: dev(:)
    integer :: k
    if (present(devices)) then
      allocate(dev(size(devices)))
      dev = int(devices, c_int)
    else
      allocate(dev(rrnn_device_count()))
      do k = 1, size(dev)
        dev(k) = int(k - 1, c_int)
      end do
    end if
    error_msg = rrnn_error_msg(rrnn_multi_create(int(size(dev), c_int), dev, this%h))
  end function multi_init

  function multi_load_netcdf(this, filename, model_id) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    character(len=*),     intent(in)    :: filename
    integer,              intent(out)   :: model_id
    character(len=128) :: error_msg
    integer(c_int) :: id(1)
    error_msg = rrnn_error_msg(rrnn_multi_model_load_netcdf(this%h, c_str(filename), id))
    model_id = int(id(1))
  end function multi_load_netcdf

  function multi_load_lw(this, band2gpt, totplnk, temp_ref_min, totplnk_delta, kdist_id) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    integer,              intent(in)    :: band2gpt(:,:)
    real(wp), contiguous, target, intent(in) :: totplnk(:,:)
    real(wp),             intent(in)    :: temp_ref_min, totplnk_delta
    integer,              intent(out)   :: kdist_id
    character(len=128) :: error_msg
    integer(c_int) :: id(1)
    integer(c_int), allocatable :: lims(:,:)
    allocate(lims(2, size(band2gpt, 2)))
    lims = int(band2gpt, c_int)
    error_msg = rrnn_error_msg(rrnn_multi_kdist_create(this%h, int(size(lims, 2), c_int), int(maxval(lims), c_int), lims, &
                  int(size(totplnk, 1), c_int), c_loc(totplnk), temp_ref_min, totplnk_delta, c_null_ptr, id))
    kdist_id = int(id(1))
  end function multi_load_lw

  function multi_load_sw(this, band2gpt, solar_source, kdist_id) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    integer,              intent(in)    :: band2gpt(:,:)
    real(wp), contiguous, target, intent(in) :: solar_source(:)
    integer,              intent(out)   :: kdist_id
    character(len=128) :: error_msg
    integer(c_int) :: id(1)
    integer(c_int), allocatable :: lims(:,:)
    allocate(lims(2, size(band2gpt, 2)))
    lims = int(band2gpt, c_int)
    error_msg = rrnn_error_msg(rrnn_multi_kdist_create(this%h, int(size(lims, 2), c_int), int(maxval(lims), c_int), lims, &
                  0_c_int, c_null_ptr, 0._wp, 1._wp, c_loc(solar_source), id))
    kdist_id = int(id(1))
  end function multi_load_sw

  function multi_lw(this, kdist_id, model_ids, play, plev, tlay, tsfc, sfc_emis, gas_desc, top_at_1, flux_up, flux_dn, &
                    tlev, n_gauss_angles) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    integer,              intent(in)    :: kdist_id, model_ids(:)
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), tsfc(:), sfc_emis(:)
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:)
    real(wp), contiguous, target, optional, intent(in) :: tlev(:,:)
    integer,                      optional, intent(in) :: n_gauss_angles
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    integer(c_int), allocatable :: ids(:)
    type(c_ptr) :: p_tlev
    integer :: nang
    allocate(ids(size(model_ids)))
    ids = int(model_ids, c_int)
    p_tlev = c_null_ptr
    if (present(tlev)) p_tlev = c_loc(tlev)
    nang = 1
    if (present(n_gauss_angles)) nang = n_gauss_angles
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_multi_lw_fluxes_host(this%h, int(kdist_id, c_int), ids, int(size(ids), c_int), &
                  int(size(play, 2), c_int), int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), int(nang, c_int), &
                  c_loc(play), c_loc(plev), c_loc(tlay), p_tlev, c_loc(tsfc), c_loc(sfc_emis), gases, int(size(gases), c_int), &
                  c_loc(flux_up), c_loc(flux_dn)))
  end function multi_lw

  function multi_sw(this, kdist_id, model_ids, play, plev, tlay, mu0, sfc_alb, gas_desc, top_at_1, flux_up, flux_dn, &
                    flux_dn_dir, tsi_scale) result(error_msg)
    class(ty_rrnn_multi), intent(inout) :: this
    integer,              intent(in)    :: kdist_id, model_ids(2)
    real(wp), contiguous, target, intent(in)  :: play(:,:), plev(:,:), tlay(:,:), mu0(:), sfc_alb(:)
    type(ty_gas_concs), target,   intent(in)  :: gas_desc
    logical,                      intent(in)  :: top_at_1
    real(wp), contiguous, target, intent(out) :: flux_up(:,:), flux_dn(:,:), flux_dn_dir(:,:)
    real(wp), contiguous, target, optional, intent(in) :: tsi_scale(:)
    character(len=128) :: error_msg
    type(rrnn_gas_t), allocatable :: gases(:)
    integer(c_int) :: ids(2)
    type(c_ptr) :: p_tsi
    ids = int(model_ids, c_int)
    p_tsi = c_null_ptr
    if (present(tsi_scale)) p_tsi = c_loc(tsi_scale)
    call host_gases(gas_desc, gases)
    error_msg = rrnn_error_msg(rrnn_multi_sw_fluxes_host(this%h, int(kdist_id, c_int), ids, int(size(play, 2), c_int), &
                  int(size(play, 1), c_int), merge(1_c_int, 0_c_int, top_at_1), c_loc(play), c_loc(plev), c_loc(tlay), &
                  c_loc(mu0), c_loc(sfc_alb), p_tsi, gases, int(size(gases), c_int), c_loc(flux_up), c_loc(flux_dn), &
                  c_loc(flux_dn_dir)))
  end function multi_sw

  subroutine multi_finalize(this)
    class(ty_rrnn_multi), intent(inout) :: this
    integer(c_int) :: rc
    if (c_associated(this%h)) rc = rrnn_multi_destroy(this%h)
    this%h = c_null_ptr
  end subroutine multi_finalize

end module mo_rrnn_drivers
