! Reference-facing Fortran veneer of librrnn_b200.so: the reference's OWN module, type and procedure names with the reference's
! argument lists, bodies = calls of the C ABI (mo_rrnn_c_binding).  A host model that today says
!
!     use mo_gas_optics_rrtmgp, only: ty_gas_optics_rrtmgp
!     use mod_network_rrtmgp,   only: rrtmgp_network_type
!     use mo_rte_lw,            only: rte_lw
!     ...
!     call neural_nets(1)%load_netcdf(file_tau); call neural_nets(2)%load_netcdf(file_pfrac)
!     error_msg = k_dist%gas_optics(p_lay, p_lev, t_lay, sfc_t, gas_concs, optical_props, source, tlev=t_lev, neural_nets=neural_nets)
!     error_msg = rte_lw(optical_props, top_at_1, source, sfc_emis_spec, fluxes, n_gauss_angles=n_quad_angles)
!
! (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446) links these modules instead of the reference's and keeps those lines.
!
! What differs from the reference, by design:
!   * the optical-property and source types own DEVICE buffers (type(c_ptr)) instead of allocatable host arrays
!     (rte/mo_optical_props.F90:98-192, rte/mo_source_functions.F90:26-43): tau / ssa / g / lay_source ... never cross PCIe.
!     `get_tau`, `get_ssa`, ... copy one back when a caller wants to look at it.
!   * profiles, boundary conditions and fluxes stay host arrays exactly as in the reference; every call uploads its inputs
!     and downloads its fluxes (the fused whole-path calls of mo_rrnn_drivers avoid the intermediate arrays altogether).
!   * ty_gas_optics_rrtmgp%load takes the spectral tables the NN path needs (band limits, totplnk, solar source), not the
!     k-distribution (rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326): the LUT branch is out of scope (DESIGN.md section 7), so
!     gas_optics without neural_nets returns an error instead of silently computing something else.
!   * real(wp) is c_float (the fork's single-precision build, rte/mo_rte_kind.F90:32).
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler): tests/test_fortran_cpu.py checks that every C function called
! here exists in mo_rrnn_c_binding with that many arguments, and that the public procedures keep the reference's argument names.

! ------------------------------------------------------------------------------------------------------------------------------
module mo_rte_kind
  use, intrinsic :: iso_c_binding, only: c_float, c_double, c_int, c_bool
  implicit none
  public
  integer, parameter :: sp = c_float, dp = c_double
  integer, parameter :: wp = sp            ! rte/mo_rte_kind.F90:32
  integer, parameter :: wl = kind(.true.)
end module mo_rte_kind

! ------------------------------------------------------------------------------------------------------------------------------
! The process-wide context (device 0 unless rrnn_select_device was called first) and a device-buffer helper.
module mo_rrnn_device
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  implicit none
  private
  public :: rrnn_ctx, rrnn_select_device, rrnn_shutdown, ty_devbuf, c_str

  type(c_ptr), save :: the_ctx = c_null_ptr
  integer,     save :: the_device = 0

  ! A device array of real(wp): n elements at p.
  type :: ty_devbuf
    type(c_ptr)       :: p = c_null_ptr
    integer(c_size_t) :: n = 0
  contains
    procedure :: resize   => devbuf_resize
    procedure :: free     => devbuf_free
    procedure :: is_alloc => devbuf_is_alloc
    procedure :: upload   => devbuf_upload       ! host (contiguous, any rank via c_loc) -> device
    procedure :: download => devbuf_download
  end type ty_devbuf

contains

  subroutine rrnn_select_device(device)
    integer, intent(in) :: device
    the_device = device
  end subroutine rrnn_select_device

  function rrnn_ctx() result(ctx)
    type(c_ptr) :: ctx
    integer(c_int) :: rc
    if (.not. c_associated(the_ctx)) then
      rc = rrnn_ctx_create(int(the_device, c_int), c_null_ptr, the_ctx)
      if (rc /= 0) then
        write (*, '(a)') "librrnn_b200: " // trim(rrnn_error_msg(rc))     ! no CUDA device: there is no CPU fallback
        error stop 1
      end if
    end if
    ctx = the_ctx
  end function rrnn_ctx

  subroutine rrnn_shutdown()
    integer(c_int) :: rc
    if (c_associated(the_ctx)) rc = rrnn_ctx_destroy(the_ctx)
    the_ctx = c_null_ptr
  end subroutine rrnn_shutdown

  function c_str(s) result(cs)
    character(len=*), intent(in) :: s
    character(kind=c_char, len=len_trim(s) + 1) :: cs
    cs = trim(s) // c_null_char
  end function c_str

  function devbuf_resize(this, n) result(rc)
    class(ty_devbuf), intent(inout) :: this
    integer(c_size_t), intent(in)   :: n
    integer(c_int) :: rc
    rc = 0
    if (this%n == n .and. c_associated(this%p)) return
    rc = rrnn_dev_free(rrnn_ctx(), this%p)
    this%p = c_null_ptr
    this%n = 0
    if (n == 0) return
    rc = rrnn_dev_malloc(rrnn_ctx(), n * c_sizeof(1.0_wp), this%p)
    if (rc == 0) this%n = n
  end function devbuf_resize

  subroutine devbuf_free(this)
    class(ty_devbuf), intent(inout) :: this
    integer(c_int) :: rc
    rc = rrnn_dev_free(rrnn_ctx(), this%p)
    this%p = c_null_ptr
    this%n = 0
  end subroutine devbuf_free

  logical function devbuf_is_alloc(this)
    class(ty_devbuf), intent(in) :: this
    devbuf_is_alloc = c_associated(this%p)
  end function devbuf_is_alloc

  function devbuf_upload(this, host, n) result(rc)
    class(ty_devbuf), intent(inout) :: this
    type(c_ptr),       intent(in)   :: host          ! c_loc of a contiguous real(wp) array
    integer(c_size_t), intent(in)   :: n
    integer(c_int) :: rc
    rc = this%resize(n)
    if (rc == 0) rc = rrnn_memcpy_h2d(rrnn_ctx(), this%p, host, n * c_sizeof(1.0_wp))
  end function devbuf_upload

  function devbuf_download(this, host, n) result(rc)
    class(ty_devbuf), intent(in)  :: this
    type(c_ptr),       intent(in) :: host
    integer(c_size_t), intent(in) :: n
    integer(c_int) :: rc
    rc = rrnn_memcpy_d2h(rrnn_ctx(), host, this%p, min(n, this%n) * c_sizeof(1.0_wp))
  end function devbuf_download

end module mo_rrnn_device

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_rte_rrtmgp_config.F90:23-62
module mo_rte_rrtmgp_config
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wl
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, c_str
  implicit none
  private
  logical(wl), protected, public :: check_extents = .false.
  logical(wl), protected, public :: check_values  = .false.
  interface rte_rrtmgp_config_checks
    module procedure rte_rrtmgp_config_checks_each, rte_rrtmgp_config_checks_all
  end interface
  public :: rte_rrtmgp_config_checks
contains
  subroutine rte_rrtmgp_config_checks_each(extents, values)
    logical(wl), intent(in) :: extents, values
    integer(c_int) :: rc
    check_extents = extents
    check_values  = values
    ! the library's own checks (gas_optics: rrtmgp/mo_gas_optics_rrtmgp.F90:287-315, 478-494) follow the same switches
    rc = rrnn_ctx_set_flag(rrnn_ctx(), c_str("check_extents"), merge(1_c_int, 0_c_int, check_extents))
    rc = rrnn_ctx_set_flag(rrnn_ctx(), c_str("check_values"),  merge(1_c_int, 0_c_int, check_values))
  end subroutine rte_rrtmgp_config_checks_each
  subroutine rte_rrtmgp_config_checks_all(do_checks)
    logical(wl), intent(in) :: do_checks
    call rte_rrtmgp_config_checks_each(do_checks, do_checks)
  end subroutine rte_rrtmgp_config_checks_all
end module mo_rte_rrtmgp_config

! ------------------------------------------------------------------------------------------------------------------------------
! neural/mod_network_rrtmgp.F90:34-122
module mod_network_rrtmgp
  use, intrinsic :: iso_c_binding
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, c_str
  implicit none
  private
  public :: rrtmgp_network_type

  type :: rrtmgp_network_type
    type(c_ptr) :: handle = c_null_ptr                 ! rrnn_model_t*: weights, biases, scaling coefficients on the device
  contains
    procedure, public, pass(self) :: load_netcdf       ! neural/mod_network_rrtmgp.F90:58-122
    procedure, public, pass(self) :: load              ! network_type%load (ASCII), neural/mod_network.F90:163-209
    procedure, public, pass(self) :: finalize => network_finalize
  end type rrtmgp_network_type

contains

  subroutine load_netcdf(self, filename)
    class(rrtmgp_network_type), intent(in out) :: self
    character(len=*),           intent(in)     :: filename
    integer(c_int) :: rc
    rc = rrnn_model_load_netcdf(rrnn_ctx(), c_str(filename), self%handle)
    if (rc /= 0) then                                   ! the reference stops in nf90 error handling
      write (*, '(a)') "load_netcdf: " // trim(rrnn_error_msg(rc))
      error stop 1
    end if
  end subroutine load_netcdf

  subroutine load(self, filename)
    class(rrtmgp_network_type), intent(in out) :: self
    character(len=*),           intent(in)     :: filename
    integer(c_int) :: rc
    ! the ASCII format has no scaling coefficients (neural/mod_network.F90:163-209): they travel in "<filename>.scaling"
    rc = rrnn_model_load_ascii(rrnn_ctx(), c_str(filename), c_str(trim(filename) // ".scaling"), self%handle)
    if (rc /= 0) then
      write (*, '(a)') "load: " // trim(rrnn_error_msg(rc))
      error stop 1
    end if
  end subroutine load

  subroutine network_finalize(self)
    class(rrtmgp_network_type), intent(in out) :: self
    integer(c_int) :: rc
    if (c_associated(self%handle)) rc = rrnn_model_destroy(self%handle)
    self%handle = c_null_ptr
  end subroutine network_finalize

end module mod_network_rrtmgp

! ------------------------------------------------------------------------------------------------------------------------------
! rrtmgp/mo_gas_concentrations.F90:50-276.  Concentrations are kept on the host as the caller gave them; to_device builds the
! rrnn_gas_t array of one gas_optics call.
module mo_gas_concentrations
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: ty_devbuf
  implicit none
  private
  public :: ty_gas_concs

  type :: conc_field
    real(wp), allocatable :: conc(:,:)             ! (1,1) scalar, (nlay,1) profile, (nlay,ncol) field
  end type conc_field

  type :: ty_gas_concs
    character(len=32), allocatable :: gas_name(:)
    type(conc_field),  allocatable :: concs(:)
    integer :: ncol = 0, nlay = 0
  contains
    procedure, public :: init
    procedure, private :: set_vmr_scalar, set_vmr_1d, set_vmr_2d
    generic,   public :: set_vmr => set_vmr_scalar, set_vmr_1d, set_vmr_2d
    procedure, public :: get_num_gases
    procedure, public :: reset
    procedure, public :: to_device                  ! (veneer) -> rrnn_gas_t(:) with device pointers
  end type ty_gas_concs

contains

  function init(this, gas_names) result(error_msg)
    class(ty_gas_concs),            intent(inout) :: this
    character(len=*), dimension(:), intent(in   ) :: gas_names
    character(len=128)                            :: error_msg
    integer :: i, j
    error_msg = ""
    do i = 1, size(gas_names)
      if (len_trim(gas_names(i)) == 0) error_msg = "ty_gas_concs%init(): must provide non-empty gas names"
      do j = i + 1, size(gas_names)
        if (lower(gas_names(i)) == lower(gas_names(j))) error_msg = "ty_gas_concs%init(): duplicate gas names aren't allowed"
      end do
    end do
    if (error_msg /= "") return
    call this%reset()
    allocate(this%gas_name(size(gas_names)), this%concs(size(gas_names)))
    do i = 1, size(gas_names)
      this%gas_name(i) = lower(gas_names(i))
    end do
  end function init

  function find_gas(this, gas) result(i)
    class(ty_gas_concs), intent(in) :: this
    character(len=*),    intent(in) :: gas
    integer :: i
    if (allocated(this%gas_name)) then
      do i = 1, size(this%gas_name)
        if (trim(this%gas_name(i)) == trim(lower(gas))) return
      end do
    end if
    i = 0
  end function find_gas

  function set_vmr_scalar(this, gas, w) result(error_msg)
    class(ty_gas_concs), intent(inout) :: this
    character(len=*),    intent(in   ) :: gas
    real(wp),            intent(in   ) :: w
    character(len=128)                 :: error_msg
    integer :: igas
    error_msg = ""
    if (w < 0._wp .or. w > 1._wp) then
      error_msg = "ty_gas_concs%set_vmr(): concentrations should be >= 0, <= 1"
      return
    end if
    igas = find_gas(this, gas)
    if (igas == 0) then
      error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
      return
    end if
    if (allocated(this%concs(igas)%conc)) deallocate(this%concs(igas)%conc)
    allocate(this%concs(igas)%conc(1, 1))
    this%concs(igas)%conc(1, 1) = w
  end function set_vmr_scalar

  function set_vmr_1d(this, gas, w) result(error_msg)
    class(ty_gas_concs),    intent(inout) :: this
    character(len=*),       intent(in   ) :: gas
    real(wp), dimension(:), intent(in   ) :: w
    character(len=128)                    :: error_msg
    integer :: igas
    error_msg = ""
    if (any(w < 0._wp .or. w > 1._wp)) error_msg = "ty_gas_concs%set_vmr: concentrations should be >= 0, <= 1"
    if (this%nlay > 0 .and. size(w) /= this%nlay) error_msg = "ty_gas_concs%set_vmr: different dimension (nlay)"
    igas = find_gas(this, gas)
    if (igas == 0) error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
    if (error_msg /= "") return
    this%nlay = size(w)
    if (allocated(this%concs(igas)%conc)) deallocate(this%concs(igas)%conc)
    allocate(this%concs(igas)%conc(this%nlay, 1))
    this%concs(igas)%conc(:, 1) = w
  end function set_vmr_1d

  function set_vmr_2d(this, gas, w) result(error_msg)
    class(ty_gas_concs),      intent(inout) :: this
    character(len=*),         intent(in   ) :: gas
    real(wp), dimension(:,:), intent(in   ) :: w          ! (nlay, ncol) in this fork
    character(len=128)                      :: error_msg
    integer :: igas
    error_msg = ""
    if (any(w < 0._wp .or. w > 1._wp)) error_msg = "ty_gas_concs%set_vmr: concentrations should be >= 0, <= 1"
    if (this%nlay > 0 .and. size(w, 1) /= this%nlay) error_msg = "ty_gas_concs%set_vmr: different dimension (nlay)"
    if (this%ncol > 0 .and. size(w, 2) /= this%ncol) error_msg = "ty_gas_concs%set_vmr: different dimension (ncol)"
    igas = find_gas(this, gas)
    if (igas == 0) error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
    if (error_msg /= "") return
    this%nlay = size(w, 1)
    this%ncol = size(w, 2)
    if (allocated(this%concs(igas)%conc)) deallocate(this%concs(igas)%conc)
    allocate(this%concs(igas)%conc(this%nlay, this%ncol))
    this%concs(igas)%conc = w
  end function set_vmr_2d

  pure function get_num_gases(this)
    class(ty_gas_concs), intent(in) :: this
    integer :: get_num_gases
    get_num_gases = 0
    if (allocated(this%gas_name)) get_num_gases = size(this%gas_name)
  end function get_num_gases

  subroutine reset(this)
    class(ty_gas_concs), intent(inout) :: this
    if (allocated(this%gas_name)) deallocate(this%gas_name)
    if (allocated(this%concs)) deallocate(this%concs)
    this%nlay = 0
    this%ncol = 0
  end subroutine reset

  ! gases(:) for one C call; bufs(:) owns the device copies of the profile / field concentrations (free them after the call)
  function to_device(this, gases, bufs) result(rc)
    class(ty_gas_concs), target,   intent(in)  :: this
    type(rrnn_gas_t), allocatable, intent(out) :: gases(:)
    type(ty_devbuf),  allocatable, intent(out) :: bufs(:)
    integer(c_int) :: rc
    integer :: i, k, n
    rc = 0
    n = this%get_num_gases()
    allocate(gases(n), bufs(n))
    do i = 1, n
      gases(i)%name = c_null_char
      do k = 1, len_trim(this%gas_name(i))
        gases(i)%name(k) = this%gas_name(i)(k:k)
      end do
      gases(i)%conc = c_null_ptr
      gases(i)%value = 0._wp
      gases(i)%ndims = 0
      if (.not. allocated(this%concs(i)%conc)) cycle                 ! never set: treated as absent (value 0)
      if (size(this%concs(i)%conc) == 1) then
        gases(i)%value = this%concs(i)%conc(1, 1)
      else
        rc = bufs(i)%upload(c_loc(this%concs(i)%conc), int(size(this%concs(i)%conc), c_size_t))
        if (rc /= 0) return
        gases(i)%conc = bufs(i)%p
        gases(i)%ndims = merge(1, 2, size(this%concs(i)%conc, 2) == 1 .and. this%ncol /= 1)
      end if
    end do
  end function to_device

  pure function lower(s) result(t)
    character(len=*), intent(in) :: s
    character(len=len(s)) :: t
    integer :: i, c
    t = s
    do i = 1, len(s)
      c = iachar(s(i:i))
      if (c >= iachar("A") .and. c <= iachar("Z")) t(i:i) = achar(c + 32)
    end do
  end function lower

end module mo_gas_concentrations

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_optical_props.F90:52-218: spectral discretisation + the array types, device-resident.
module mo_optical_props
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  implicit none
  private
  public :: ty_optical_props, ty_optical_props_arry, ty_optical_props_1scl, ty_optical_props_2str

  type :: ty_optical_props
    integer,  allocatable :: band2gpt(:,:)            ! (2, nband), 1-based inclusive
    real(wp), allocatable :: band_lims_wvn(:,:)
    type(c_ptr) :: kd = c_null_ptr                    ! rrnn_kdist_t* carrying the same band limits on the device
    logical     :: owns_kd = .false.
    character(len=32) :: name = ""
  contains
    procedure, public :: init => init_base
    procedure, public :: is_initialized => is_initialized_base
    procedure, public :: finalize_base
    procedure, public :: get_nband
    procedure, public :: get_ngpt
    procedure, public :: get_band_lims_gpoint
    procedure, public :: get_band_lims_wavenumber
    procedure, public :: set_name
    procedure, public :: get_name
  end type ty_optical_props

  type, extends(ty_optical_props), abstract :: ty_optical_props_arry
    type(ty_devbuf) :: tau                            ! (ngpt, nlay, ncol) on the device
    integer :: ncol = 0, nlay = 0
  contains
    procedure, public :: get_ncol
    procedure, public :: get_nlay
    procedure, public :: get_tau                      ! (veneer) device -> host copy
    procedure, public :: increment                    ! rte/mo_optical_props.F90:714-893 (by-band increments)
    procedure(delta_scale_abstract), deferred, public :: delta_scale
  end type ty_optical_props_arry

  abstract interface
    function delta_scale_abstract(this, for) result(err_message)
      import :: ty_optical_props_arry, wp
      class(ty_optical_props_arry), intent(inout) :: this
      real(wp), dimension(:,:,:), optional, intent(in) :: for
      character(len=128) :: err_message
    end function delta_scale_abstract
  end interface

  type, extends(ty_optical_props_arry) :: ty_optical_props_1scl
  contains
    procedure, public :: delta_scale => delta_scale_1scl
    procedure, private :: alloc_only_1scl, init_and_alloc_1scl
    generic,   public :: alloc_1scl => alloc_only_1scl, init_and_alloc_1scl
    procedure, public :: finalize => finalize_1scl
  end type ty_optical_props_1scl

  type, extends(ty_optical_props_arry) :: ty_optical_props_2str
    type(ty_devbuf) :: ssa, g                         ! g stays unallocated on the NN path (identically 0,
                                                      ! rrtmgp/mo_gas_optics_rrtmgp.F90:560-567) until something needs it
  contains
    procedure, public :: delta_scale => delta_scale_2str
    procedure, private :: alloc_only_2str, init_and_alloc_2str
    generic,   public :: alloc_2str => alloc_only_2str, init_and_alloc_2str
    procedure, public :: get_ssa, get_g
    procedure, public :: finalize => finalize_2str
  end type ty_optical_props_2str

contains

  function init_base(this, band_lims_wvn, band_lims_gpt, name) result(err_message)
    class(ty_optical_props),  intent(inout) :: this
    real(wp), dimension(:,:), intent(in   ) :: band_lims_wvn
    integer,  dimension(:,:), optional, intent(in) :: band_lims_gpt
    character(len=*),         optional, intent(in) :: name
    character(len=128) :: err_message
    integer :: ib
    integer(c_int), allocatable :: lims(:,:)
    err_message = ""
    if (size(band_lims_wvn, 1) /= 2) err_message = "optical_props%init(): band_lims_wvn 1st dim should be 2"
    if (any(band_lims_wvn < 0._wp)) err_message = "optical_props%init(): band_lims_wvn has values <  0., respectively"
    if (err_message /= "") return
    call this%finalize_base()
    allocate(this%band_lims_wvn(2, size(band_lims_wvn, 2)), this%band2gpt(2, size(band_lims_wvn, 2)))
    this%band_lims_wvn = band_lims_wvn
    if (present(band_lims_gpt)) then
      this%band2gpt = band_lims_gpt
    else
      do ib = 1, size(band_lims_wvn, 2)
        this%band2gpt(:, ib) = ib
      end do
    end if
    if (present(name)) this%name = name
    allocate(lims(2, size(this%band2gpt, 2)))
    lims = int(this%band2gpt, c_int)
    err_message = rrnn_error_msg(rrnn_kdist_create(rrnn_ctx(), int(size(lims, 2), c_int), int(maxval(lims), c_int), lims, &
                                 0_c_int, c_null_ptr, 0._wp, 1._wp, c_null_ptr, this%kd))
    this%owns_kd = err_message == ""
  end function init_base

  pure function is_initialized_base(this)
    class(ty_optical_props), intent(in) :: this
    logical :: is_initialized_base
    is_initialized_base = allocated(this%band2gpt)
  end function is_initialized_base

  subroutine finalize_base(this)
    class(ty_optical_props), intent(inout) :: this
    integer(c_int) :: rc
    if (allocated(this%band2gpt)) deallocate(this%band2gpt)
    if (allocated(this%band_lims_wvn)) deallocate(this%band_lims_wvn)
    if (this%owns_kd .and. c_associated(this%kd)) rc = rrnn_kdist_destroy(this%kd)
    this%kd = c_null_ptr
    this%owns_kd = .false.
    this%name = ""
  end subroutine finalize_base

  pure function get_nband(this)
    class(ty_optical_props), intent(in) :: this
    integer :: get_nband
    get_nband = 0
    if (allocated(this%band2gpt)) get_nband = size(this%band2gpt, 2)
  end function get_nband

  pure function get_ngpt(this)
    class(ty_optical_props), intent(in) :: this
    integer :: get_ngpt
    get_ngpt = 0
    if (allocated(this%band2gpt)) get_ngpt = maxval(this%band2gpt)
  end function get_ngpt

  pure function get_band_lims_gpoint(this)
    class(ty_optical_props), intent(in) :: this
    integer, dimension(size(this%band2gpt, 1), size(this%band2gpt, 2)) :: get_band_lims_gpoint
    get_band_lims_gpoint = this%band2gpt
  end function get_band_lims_gpoint

  pure function get_band_lims_wavenumber(this)
    class(ty_optical_props), intent(in) :: this
    real(wp), dimension(size(this%band_lims_wvn, 1), size(this%band_lims_wvn, 2)) :: get_band_lims_wavenumber
    get_band_lims_wavenumber = this%band_lims_wvn
  end function get_band_lims_wavenumber

  subroutine set_name(this, name)
    class(ty_optical_props), intent(inout) :: this
    character(len=*),        intent(in)    :: name
    this%name = name
  end subroutine set_name

  function get_name(this)
    class(ty_optical_props), intent(in) :: this
    character(len=32) :: get_name
    get_name = this%name
  end function get_name

  pure function get_ncol(this)
    class(ty_optical_props_arry), intent(in) :: this
    integer :: get_ncol
    get_ncol = this%ncol
  end function get_ncol

  pure function get_nlay(this)
    class(ty_optical_props_arry), intent(in) :: this
    integer :: get_nlay
    get_nlay = this%nlay
  end function get_nlay

  ! the spectral discretisation of `spectral_desc` is shared, not copied: the device handle stays owned by its creator
  subroutine share_spectral(this, spectral_desc)
    class(ty_optical_props), intent(inout) :: this
    class(ty_optical_props), intent(in)    :: spectral_desc
    call this%finalize_base()
    allocate(this%band2gpt(2, spectral_desc%get_nband()), this%band_lims_wvn(2, spectral_desc%get_nband()))
    this%band2gpt = spectral_desc%band2gpt
    this%band_lims_wvn = spectral_desc%band_lims_wvn
    this%kd = spectral_desc%kd
    this%owns_kd = .false.
  end subroutine share_spectral

  function alloc_only_1scl(this, ncol, nlay) result(err_message)
    class(ty_optical_props_1scl), intent(inout) :: this
    integer,                      intent(in)    :: ncol, nlay
    character(len=128) :: err_message
    err_message = ""
    if (.not. this%is_initialized()) then
      err_message = "optical_props%alloc: spectral discretization hasn't been provided"
    else if (any([ncol, nlay] <= 0)) then
      err_message = "optical_props%alloc: must provide positive extents for ncol, nlay"
    else
      this%ncol = ncol
      this%nlay = nlay
      err_message = rrnn_error_msg(this%tau%resize(int(this%get_ngpt(), c_size_t) * nlay * ncol))
    end if
  end function alloc_only_1scl

  function init_and_alloc_1scl(this, ncol, nlay, spectral_desc, name) result(err_message)
    class(ty_optical_props_1scl), intent(inout) :: this
    integer,                      intent(in)    :: ncol, nlay
    class(ty_optical_props),      intent(in)    :: spectral_desc
    character(len=*), optional,   intent(in)    :: name
    character(len=128) :: err_message
    if (.not. spectral_desc%is_initialized()) then
      err_message = "optical_props%alloc: spectral discretization hasn't been provided"
      return
    end if
    call share_spectral(this, spectral_desc)
    if (present(name)) this%name = name
    err_message = this%alloc_1scl(ncol, nlay)
  end function init_and_alloc_1scl

  function alloc_only_2str(this, ncol, nlay) result(err_message)
    class(ty_optical_props_2str), intent(inout) :: this
    integer,                      intent(in)    :: ncol, nlay
    character(len=128) :: err_message
    integer(c_size_t) :: n
    err_message = ""
    if (.not. this%is_initialized()) then
      err_message = "optical_props%alloc: spectral discretization hasn't been provided"
    else if (any([ncol, nlay] <= 0)) then
      err_message = "optical_props%alloc: must provide positive extents for ncol, nlay"
    else
      this%ncol = ncol
      this%nlay = nlay
      n = int(this%get_ngpt(), c_size_t) * nlay * ncol
      err_message = rrnn_error_msg(this%tau%resize(n))
      if (err_message == "") err_message = rrnn_error_msg(this%ssa%resize(n))
      call this%g%free()
    end if
  end function alloc_only_2str

  function init_and_alloc_2str(this, ncol, nlay, spectral_desc, name) result(err_message)
    class(ty_optical_props_2str), intent(inout) :: this
    integer,                      intent(in)    :: ncol, nlay
    class(ty_optical_props),      intent(in)    :: spectral_desc
    character(len=*), optional,   intent(in)    :: name
    character(len=128) :: err_message
    if (.not. spectral_desc%is_initialized()) then
      err_message = "optical_props%alloc: spectral discretization hasn't been provided"
      return
    end if
    call share_spectral(this, spectral_desc)
    if (present(name)) this%name = name
    err_message = this%alloc_2str(ncol, nlay)
  end function init_and_alloc_2str

  subroutine finalize_1scl(this)
    class(ty_optical_props_1scl), intent(inout) :: this
    call this%tau%free()
    call this%finalize_base()
  end subroutine finalize_1scl

  subroutine finalize_2str(this)
    class(ty_optical_props_2str), intent(inout) :: this
    call this%tau%free()
    call this%ssa%free()
    call this%g%free()
    call this%finalize_base()
  end subroutine finalize_2str

  function delta_scale_1scl(this, for) result(err_message)            ! rte/mo_optical_props.F90:438-447: nothing to do
    class(ty_optical_props_1scl), intent(inout) :: this
    real(wp), dimension(:,:,:), optional, intent(in) :: for
    character(len=128) :: err_message
    err_message = ""
  end function delta_scale_1scl

  function delta_scale_2str(this, for) result(err_message)            ! :449-477 -> delta_scale_2str_k
    class(ty_optical_props_2str), intent(inout) :: this
    real(wp), dimension(:,:,:), optional, intent(in) :: for
    character(len=128) :: err_message
    err_message = ""
    if (present(for)) then
      err_message = "delta_scale: user-provided forward-scattering fraction is not supported by librrnn_b200"
      return
    end if
    if (.not. this%g%is_alloc()) return                                 ! g == 0: f = 0, tau and ssa unchanged
    err_message = rrnn_error_msg(rrnn_delta_scale_2str(rrnn_ctx(), this%tau%n, this%tau%p, this%ssa%p, this%g%p))
  end function delta_scale_2str

  ! op_io = op_io + op_in, op_in on the same grid but by BAND (cloud optics): inc_*_bybnd kernels
  function increment(op_in, op_io) result(err_message)
    class(ty_optical_props_arry), intent(in   ) :: op_in
    class(ty_optical_props_arry), intent(inout) :: op_io
    character(len=128) :: err_message
    integer(c_int) :: rc
    err_message = ""
    if (op_in%ncol /= op_io%ncol .or. op_in%nlay /= op_io%nlay) then
      err_message = "ty_optical_props%increment: optical properties objects have different ncol and/or nlay"
      return
    end if
    if (op_in%get_ngpt() /= op_io%get_nband()) then
      err_message = "ty_optical_props%increment: optical properties objects have incompatible g-point structures"
      return
    end if
    select type (op_io)
    class is (ty_optical_props_1scl)
      rc = rrnn_increment_1scl_bybnd(rrnn_ctx(), op_io%kd, int(op_io%nlay, c_int), int(op_io%ncol, c_int), op_io%tau%p, op_in%tau%p)
      err_message = rrnn_error_msg(rc)
    class is (ty_optical_props_2str)
      select type (op_in)
      class is (ty_optical_props_2str)
        if (.not. op_io%g%is_alloc()) then                             ! materialise g == 0 before it is incremented
          rc = op_io%g%resize(op_io%tau%n)
          if (rc == 0) rc = rrnn_fill_zero(op_io%g)
        end if
        rc = rrnn_increment_2str_bybnd(rrnn_ctx(), op_io%kd, int(op_io%nlay, c_int), int(op_io%ncol, c_int), op_io%tau%p, &
                                       op_io%ssa%p, op_io%g%p, op_in%tau%p, op_in%ssa%p, op_in%g%p)
        err_message = rrnn_error_msg(rc)
      class default
        err_message = "ty_optical_props%increment: only 2str-by-2str and 1scl-by-1scl increments are provided"
      end select
    class default
      err_message = "ty_optical_props%increment: unknown optical properties type"
    end select
  end function increment

  function rrnn_fill_zero(buf) result(rc)
    type(ty_devbuf), intent(inout) :: buf
    integer(c_int) :: rc
    real(wp), allocatable, target :: z(:)
    allocate(z(buf%n))
    z = 0._wp
    rc = buf%upload(c_loc(z), buf%n)
  end function rrnn_fill_zero

  function get_tau(this, tau) result(err_message)
    class(ty_optical_props_arry), intent(in)  :: this
    real(wp), contiguous, target, intent(out) :: tau(:,:,:)
    character(len=128) :: err_message
    err_message = rrnn_error_msg(this%tau%download(c_loc(tau), int(size(tau), c_size_t)))
  end function get_tau

  function get_ssa(this, ssa) result(err_message)
    class(ty_optical_props_2str), intent(in)  :: this
    real(wp), contiguous, target, intent(out) :: ssa(:,:,:)
    character(len=128) :: err_message
    err_message = rrnn_error_msg(this%ssa%download(c_loc(ssa), int(size(ssa), c_size_t)))
  end function get_ssa

  function get_g(this, g) result(err_message)
    class(ty_optical_props_2str), intent(in)  :: this
    real(wp), contiguous, target, intent(out) :: g(:,:,:)
    character(len=128) :: err_message
    err_message = ""
    g = 0._wp
    if (this%g%is_alloc()) err_message = rrnn_error_msg(this%g%download(c_loc(g), int(size(g), c_size_t)))
  end function get_g

end module mo_optical_props

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_source_functions.F90:26-43
module mo_source_functions
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: ty_devbuf
  use mo_optical_props, only: ty_optical_props
  implicit none
  private
  public :: ty_source_func_lw

  type, extends(ty_optical_props) :: ty_source_func_lw
    type(ty_devbuf) :: lay_source, lev_source        ! (ngpt, nlay, ncol), (ngpt, nlay+1, ncol)
    type(ty_devbuf) :: sfc_source, sfc_source_Jac    ! (ngpt, ncol)
    integer :: ncol = 0, nlay = 0
  contains
    procedure, private :: alloc_lw, copy_and_alloc_lw
    generic,   public  :: alloc => alloc_lw, copy_and_alloc_lw
    procedure, public  :: is_allocated => is_allocated_lw
    procedure, public  :: finalize => finalize_lw
    procedure, public  :: get_ncol => get_ncol_lw
    procedure, public  :: get_nlay => get_nlay_lw
  end type ty_source_func_lw

contains

  pure function is_allocated_lw(this)
    class(ty_source_func_lw), intent(in) :: this
    logical :: is_allocated_lw
    is_allocated_lw = this%is_initialized() .and. c_associated(this%sfc_source%p)
  end function is_allocated_lw

  function alloc_lw(this, ncol, nlay) result(err_message)
    class(ty_source_func_lw), intent(inout) :: this
    integer,                  intent(in)    :: ncol, nlay
    character(len=128) :: err_message
    integer(c_size_t) :: ngpt
    integer(c_int) :: rc
    err_message = ""
    if (.not. this%is_initialized()) err_message = "source_func_lw%alloc: not initialized so can't allocate"
    if (any([ncol, nlay] <= 0)) err_message = "source_func_lw%alloc: must provide positive extents for ncol, nlay"
    if (err_message /= "") return
    this%ncol = ncol
    this%nlay = nlay
    ngpt = int(this%get_ngpt(), c_size_t)
    rc = this%lay_source%resize(ngpt * nlay * ncol)
    if (rc == 0) rc = this%lev_source%resize(ngpt * (nlay + 1) * ncol)
    if (rc == 0) rc = this%sfc_source%resize(ngpt * ncol)
    if (rc == 0) rc = this%sfc_source_Jac%resize(ngpt * ncol)
    err_message = rrnn_error_msg(rc)
  end function alloc_lw

  function copy_and_alloc_lw(this, ncol, nlay, spectral_desc) result(err_message)
    class(ty_source_func_lw), intent(inout) :: this
    integer,                  intent(in)    :: ncol, nlay
    class(ty_optical_props),  intent(in)    :: spectral_desc
    character(len=128) :: err_message
    if (.not. spectral_desc%is_initialized()) then
      err_message = "source_func_lw%alloc: spectral_desc not initialized"
      return
    end if
    call this%finalize_base()
    allocate(this%band2gpt(2, spectral_desc%get_nband()), this%band_lims_wvn(2, spectral_desc%get_nband()))
    this%band2gpt = spectral_desc%band2gpt
    this%band_lims_wvn = spectral_desc%band_lims_wvn
    this%kd = spectral_desc%kd
    this%owns_kd = .false.
    err_message = this%alloc(ncol, nlay)
  end function copy_and_alloc_lw

  subroutine finalize_lw(this)
    class(ty_source_func_lw), intent(inout) :: this
    call this%lay_source%free()
    call this%lev_source%free()
    call this%sfc_source%free()
    call this%sfc_source_Jac%free()
    call this%finalize_base()
  end subroutine finalize_lw

  pure function get_ncol_lw(this)
    class(ty_source_func_lw), intent(in) :: this
    integer :: get_ncol_lw
    get_ncol_lw = this%ncol
  end function get_ncol_lw

  pure function get_nlay_lw(this)
    class(ty_source_func_lw), intent(in) :: this
    integer :: get_nlay_lw
    get_nlay_lw = this%nlay
  end function get_nlay_lw

end module mo_source_functions

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_fluxes.F90:46-67: host pointers, results written straight into the caller's memory.
module mo_fluxes
  use mo_rte_kind, only: wp
  implicit none
  private
  public :: ty_fluxes, ty_fluxes_broadband, ty_fluxes_flexible

  type, abstract :: ty_fluxes
  end type ty_fluxes

  type, extends(ty_fluxes) :: ty_fluxes_broadband
    real(wp), dimension(:,:), contiguous, pointer :: flux_up => NULL(), flux_dn => NULL()       ! (nlay+1, ncol)
    real(wp), dimension(:,:), contiguous, pointer :: flux_net => NULL()                          ! down - up
    real(wp), dimension(:,:), contiguous, pointer :: flux_dn_dir => NULL()
  contains
    procedure, public :: are_desired => are_desired_broadband
  end type ty_fluxes_broadband

  type, extends(ty_fluxes_broadband) :: ty_fluxes_flexible
    real(wp), dimension(:,:,:), contiguous, pointer :: gpt_flux_up => NULL(), gpt_flux_dn => NULL()   ! (ngpt, nlay+1, ncol)
    real(wp), dimension(:,:,:), contiguous, pointer :: gpt_flux_net => NULL()
    real(wp), dimension(:,:,:), contiguous, pointer :: gpt_flux_dn_dir => NULL()
    real(wp), dimension(:,:,:), contiguous, pointer :: gpt_flux_up_Jac => NULL()
  contains
    procedure, public :: are_desired_gpt
  end type ty_fluxes_flexible

contains

  function are_desired_broadband(this)
    class(ty_fluxes_broadband), intent(in) :: this
    logical :: are_desired_broadband
    are_desired_broadband = any([associated(this%flux_up), associated(this%flux_dn), associated(this%flux_dn_dir), &
                                 associated(this%flux_net)])
  end function are_desired_broadband

  function are_desired_gpt(this)
    class(ty_fluxes_flexible), intent(in) :: this
    logical :: are_desired_gpt
    are_desired_gpt = any([associated(this%gpt_flux_up), associated(this%gpt_flux_dn), associated(this%gpt_flux_dn_dir), &
                           associated(this%gpt_flux_net)])
  end function are_desired_gpt

end module mo_fluxes

! ------------------------------------------------------------------------------------------------------------------------------
! rrtmgp/mo_gas_optics_rrtmgp.F90: the NN branches of gas_optics (gas_optics_int :239-428, gas_optics_ext :433-602), set_tsi.
module mo_gas_optics_rrtmgp
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  use mo_optical_props, only: ty_optical_props, ty_optical_props_arry, ty_optical_props_1scl, ty_optical_props_2str
  use mo_source_functions, only: ty_source_func_lw
  use mo_gas_concentrations, only: ty_gas_concs
  use mod_network_rrtmgp, only: rrtmgp_network_type
  implicit none
  private
  public :: ty_gas_optics_rrtmgp

  type, extends(ty_optical_props) :: ty_gas_optics_rrtmgp
    logical :: has_planck = .false., has_solar = .false.
    real(wp) :: press_ref_min = 0._wp, press_ref_max = huge(1._wp), temp_ref_min = 0._wp, temp_ref_max = huge(1._wp)
  contains
    procedure, private :: load_int, load_ext
    generic,   public  :: load => load_int, load_ext
    procedure, private :: gas_optics_int, gas_optics_ext
    generic,   public  :: gas_optics => gas_optics_int, gas_optics_ext
    procedure, public  :: source_is_internal
    procedure, public  :: source_is_external
    procedure, public  :: set_tsi
    procedure, public  :: set_solar_variability
    procedure, public  :: get_press_min, get_press_max, get_temp_min, get_temp_max
    procedure, public  :: finalize => gas_optics_finalize
  end type ty_gas_optics_rrtmgp

contains

  ! Internal sources (longwave): band structure + the Planck table totplnk (nPlanckTemp, nbnd) on temp_ref_min + k*totplnk_delta
  function load_int(this, band_lims_wvn, band2gpt, totplnk, temp_ref_min, totplnk_delta, press_ref, temp_ref) result(err_message)
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    real(wp), dimension(:,:),     intent(in)   :: band_lims_wvn
    integer,  dimension(:,:),     intent(in)   :: band2gpt
    real(wp), dimension(:,:), contiguous, target, intent(in) :: totplnk
    real(wp),                     intent(in)   :: temp_ref_min, totplnk_delta
    real(wp), dimension(:), optional, intent(in) :: press_ref, temp_ref       ! enable the range checks of :478-494
    character(len=128) :: err_message
    integer(c_int), allocatable :: lims(:,:)
    call this%finalize()
    allocate(this%band_lims_wvn(2, size(band2gpt, 2)), this%band2gpt(2, size(band2gpt, 2)), lims(2, size(band2gpt, 2)))
    this%band_lims_wvn = band_lims_wvn
    this%band2gpt = band2gpt
    lims = int(band2gpt, c_int)
    err_message = rrnn_error_msg(rrnn_kdist_create(rrnn_ctx(), int(size(lims, 2), c_int), int(maxval(lims), c_int), lims, &
                                 int(size(totplnk, 1), c_int), c_loc(totplnk), temp_ref_min, totplnk_delta, c_null_ptr, this%kd))
    this%owns_kd = err_message == ""
    this%has_planck = .true.
    call set_ranges(this, press_ref, temp_ref)
  end function load_int

  ! External sources (shortwave): band structure + solar_source (ngpt)
  function load_ext(this, band_lims_wvn, band2gpt, solar_source, press_ref, temp_ref) result(err_message)
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    real(wp), dimension(:,:),     intent(in)   :: band_lims_wvn
    integer,  dimension(:,:),     intent(in)   :: band2gpt
    real(wp), dimension(:), contiguous, target, intent(in) :: solar_source
    real(wp), dimension(:), optional, intent(in) :: press_ref, temp_ref
    character(len=128) :: err_message
    integer(c_int), allocatable :: lims(:,:)
    call this%finalize()
    allocate(this%band_lims_wvn(2, size(band2gpt, 2)), this%band2gpt(2, size(band2gpt, 2)), lims(2, size(band2gpt, 2)))
    this%band_lims_wvn = band_lims_wvn
    this%band2gpt = band2gpt
    lims = int(band2gpt, c_int)
    err_message = rrnn_error_msg(rrnn_kdist_create(rrnn_ctx(), int(size(lims, 2), c_int), int(maxval(lims), c_int), lims, &
                                 0_c_int, c_null_ptr, 0._wp, 1._wp, c_loc(solar_source), this%kd))
    this%owns_kd = err_message == ""
    this%has_solar = .true.
    call set_ranges(this, press_ref, temp_ref)
  end function load_ext

  subroutine set_ranges(this, press_ref, temp_ref)
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    real(wp), dimension(:), optional, intent(in) :: press_ref, temp_ref
    if (present(press_ref)) then
      this%press_ref_min = minval(press_ref)
      this%press_ref_max = maxval(press_ref)
    end if
    if (present(temp_ref)) then
      this%temp_ref_min = minval(temp_ref)
      this%temp_ref_max = maxval(temp_ref)
    end if
  end subroutine set_ranges

  subroutine gas_optics_finalize(this)
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    call this%finalize_base()
    this%has_planck = .false.
    this%has_solar = .false.
  end subroutine gas_optics_finalize

  pure function source_is_internal(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    logical :: source_is_internal
    source_is_internal = this%has_planck
  end function source_is_internal

  pure function source_is_external(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    logical :: source_is_external
    source_is_external = this%has_solar
  end function source_is_external

  pure function get_press_min(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp) :: get_press_min
    get_press_min = this%press_ref_min
  end function get_press_min
  pure function get_press_max(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp) :: get_press_max
    get_press_max = this%press_ref_max
  end function get_press_max
  pure function get_temp_min(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp) :: get_temp_min
    get_temp_min = this%temp_ref_min
  end function get_temp_min
  pure function get_temp_max(this)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp) :: get_temp_max
    get_temp_max = this%temp_ref_max
  end function get_temp_max

  function set_tsi(this, tsi) result(error_msg)                       ! :1097-1120
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    real(wp),                    intent(in)    :: tsi
    character(len=128) :: error_msg
    if (tsi < 0._wp) then
      error_msg = "set_tsi: tsi must be > 0"
    else
      error_msg = rrnn_error_msg(rrnn_kdist_set_tsi(this%kd, tsi))
    end if
  end function set_tsi

  function set_solar_variability(this, mg_index, sb_index, tsi) result(error_msg)      ! :1058-1095
    class(ty_gas_optics_rrtmgp), intent(inout) :: this
    real(wp),                    intent(in)    :: mg_index, sb_index
    real(wp), optional,          intent(in)    :: tsi
    character(len=128) :: error_msg
    if (present(tsi)) then
      error_msg = rrnn_error_msg(rrnn_kdist_set_solar_variability(this%kd, mg_index, sb_index, 1_c_int, tsi))
    else
      error_msg = rrnn_error_msg(rrnn_kdist_set_solar_variability(this%kd, mg_index, sb_index, 0_c_int, 0._wp))
    end if
  end function set_solar_variability

  function range_check(this, play, plev, tlay) result(error_msg)      ! :478-494 (check_values)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp), dimension(:,:),    intent(in) :: play, plev, tlay
    character(len=128) :: error_msg
    error_msg = ""
    if (any(play < this%press_ref_min) .or. any(play > this%press_ref_max)) error_msg = "gas_optics(): array play has values outside range"
    if (any(plev < 0._wp)) error_msg = "gas_optics(): array plev has values outside range"
    if (any(tlay < this%temp_ref_min) .or. any(tlay > this%temp_ref_max)) error_msg = "gas_optics(): array tlay has values outside range"
  end function range_check

  ! ---- gas_optics_int: rrtmgp/mo_gas_optics_rrtmgp.F90:239-428, the reference's argument list
  function gas_optics_int(this,                             &
                          play, plev, tlay, tsfc, gas_desc, &
                          optical_props, sources,           &
                          col_dry, tlev, neural_nets        &
                          ) result(error_msg)
    class(ty_gas_optics_rrtmgp), intent(in) :: this
    real(wp), dimension(:,:), contiguous, target, intent(in) :: play, plev, tlay       ! (nlay,ncol), (nlay+1,ncol), (nlay,ncol)
    real(wp), dimension(:),   contiguous, target, intent(in) :: tsfc                   ! (ncol)
    type(ty_gas_concs),       intent(in   ) :: gas_desc
    class(ty_optical_props_arry), intent(inout) :: optical_props
    class(ty_source_func_lw),     intent(inout) :: sources
    character(len=128)                      :: error_msg
    real(wp), dimension(:,:), contiguous, intent(in), optional, target :: col_dry, tlev
    type(rrtmgp_network_type), dimension(:), intent(in), optional :: neural_nets
    integer :: ncol, nlay, i
    integer(c_int) :: rc
    type(ty_devbuf) :: d_play, d_plev, d_tlay, d_tsfc, d_tlev
    type(rrnn_gas_t), allocatable :: gases(:)
    type(ty_devbuf),  allocatable :: gas_bufs(:)
    type(c_ptr) :: models(2), p_tlev

    nlay = size(play, dim=1)
    ncol = size(play, dim=2)
    error_msg = ""
    if (.not. present(neural_nets)) then
      error_msg = "gas_optics(): librrnn_b200 provides the neural-network branch only (pass neural_nets)"
    else if (size(neural_nets) < 1 .or. size(neural_nets) > 2) then
      error_msg = "gas_optics(): neural_nets must hold one ('both') or two (absorption, Planck fraction) models"
    else if (present(col_dry)) then
      error_msg = "gas_optics(): col_dry is computed on the device from plev and h2o (get_col_dry); do not pass it"
    else if (size(plev, 1) /= nlay + 1 .or. size(plev, 2) /= ncol .or. size(tlay, 1) /= nlay .or. size(tlay, 2) /= ncol) then
      error_msg = "gas-optics(): array play, plev or tlay has wrong size"
    else if (size(tsfc) /= ncol) then
      error_msg = "gas_optics(): array tsfc has wrong size"
    else if (optical_props%get_ncol() /= ncol .or. optical_props%get_nlay() /= nlay .or. &
             optical_props%get_ngpt() /= this%get_ngpt()) then
      error_msg = "gas_optics(): optical properties have the wrong extents"
    else if (sources%get_ncol() /= ncol .or. sources%get_nlay() /= nlay .or. sources%get_ngpt() /= this%get_ngpt()) then
      error_msg = "gas_optics%gas_optics: source function arrays inconsistently sized"
    end if
    if (error_msg == "") error_msg = range_check(this, play, plev, tlay)
    if (error_msg /= "") return

    rc = d_play%upload(c_loc(play), int(size(play), c_size_t))
    if (rc == 0) rc = d_plev%upload(c_loc(plev), int(size(plev), c_size_t))
    if (rc == 0) rc = d_tlay%upload(c_loc(tlay), int(size(tlay), c_size_t))
    if (rc == 0) rc = d_tsfc%upload(c_loc(tsfc), int(size(tsfc), c_size_t))
    p_tlev = c_null_ptr
    if (rc == 0 .and. present(tlev)) then
      rc = d_tlev%upload(c_loc(tlev), int(size(tlev), c_size_t))
      p_tlev = d_tlev%p
    end if
    if (rc == 0) rc = gas_desc%to_device(gases, gas_bufs)
    models = c_null_ptr
    do i = 1, size(neural_nets)
      models(i) = neural_nets(i)%handle
    end do
    if (rc == 0) rc = rrnn_gas_optics_lw(rrnn_ctx(), this%kd, models, int(size(neural_nets), c_int), int(ncol, c_int), &
                                         int(nlay, c_int), d_play%p, d_plev%p, d_tlay%p, d_tsfc%p, gases, int(size(gases), c_int), &
                                         p_tlev, optical_props%tau%p, sources%lay_source%p, sources%lev_source%p, &
                                         sources%sfc_source%p, sources%sfc_source_Jac%p)
    error_msg = rrnn_error_msg(rc)
    call d_play%free(); call d_plev%free(); call d_tlay%free(); call d_tsfc%free(); call d_tlev%free()
    if (allocated(gas_bufs)) then
      do i = 1, size(gas_bufs)
        call gas_bufs(i)%free()
      end do
    end if
  end function gas_optics_int

  ! ---- gas_optics_ext: rrtmgp/mo_gas_optics_rrtmgp.F90:433-602, the reference's argument list
  function gas_optics_ext(this,                         &
                          play, plev, tlay, gas_desc,   &
                          optical_props, toa_src,       &
                          col_dry, neural_nets          &
                          ) result(error_msg)
    class(ty_gas_optics_rrtmgp),  intent(in) :: this
    real(wp), dimension(:,:), contiguous, target, intent(in) :: play, plev, tlay
    type(ty_gas_concs),           intent(in) :: gas_desc
    class(ty_optical_props_arry), intent(inout) :: optical_props
    real(wp), dimension(:,:), contiguous, target, intent(out) :: toa_src              ! (ngpt, ncol)
    character(len=128)                      :: error_msg
    real(wp), dimension(:,:), contiguous, intent(in), optional, target :: col_dry
    type(rrtmgp_network_type), dimension(2), intent(in), optional :: neural_nets     ! absorption model, Rayleigh model
    integer :: ncol, nlay, i
    integer(c_int) :: rc
    type(ty_devbuf) :: d_play, d_plev, d_tlay, d_toa
    type(rrnn_gas_t), allocatable :: gases(:)
    type(ty_devbuf),  allocatable :: gas_bufs(:)
    type(c_ptr) :: models(2), p_ssa

    nlay = size(play, dim=1)
    ncol = size(play, dim=2)
    error_msg = ""
    if (.not. present(neural_nets)) then
      error_msg = "gas_optics(): librrnn_b200 provides the neural-network branch only (pass neural_nets)"
    else if (present(col_dry)) then
      error_msg = "gas_optics(): col_dry is computed on the device from plev and h2o (get_col_dry); do not pass it"
    else if (size(plev, 1) /= nlay + 1 .or. size(plev, 2) /= ncol .or. size(tlay, 1) /= nlay .or. size(tlay, 2) /= ncol) then
      error_msg = "gas-optics(): array play, plev or tlay has wrong size"
    else if (size(toa_src, 1) /= this%get_ngpt() .or. size(toa_src, 2) /= ncol) then
      error_msg = "gas_optics(): array toa_src has wrong size"
    else if (optical_props%get_ncol() /= ncol .or. optical_props%get_nlay() /= nlay .or. &
             optical_props%get_ngpt() /= this%get_ngpt()) then
      error_msg = "gas_optics(): optical properties have the wrong extents"
    end if
    if (error_msg == "") error_msg = range_check(this, play, plev, tlay)
    if (error_msg /= "") return

    rc = d_play%upload(c_loc(play), int(size(play), c_size_t))
    if (rc == 0) rc = d_plev%upload(c_loc(plev), int(size(plev), c_size_t))
    if (rc == 0) rc = d_tlay%upload(c_loc(tlay), int(size(tlay), c_size_t))
    if (rc == 0) rc = d_toa%resize(int(size(toa_src), c_size_t))
    if (rc == 0) rc = gas_desc%to_device(gases, gas_bufs)
    models(1) = neural_nets(1)%handle
    models(2) = neural_nets(2)%handle
    p_ssa = c_null_ptr                                                ! 1scl request: absorption optical depth only
    select type (optical_props)
    class is (ty_optical_props_2str)
      p_ssa = optical_props%ssa%p
    end select
    if (rc == 0) rc = rrnn_gas_optics_sw(rrnn_ctx(), this%kd, models, int(ncol, c_int), int(nlay, c_int), d_play%p, d_plev%p, &
                                         d_tlay%p, gases, int(size(gases), c_int), optical_props%tau%p, p_ssa, c_null_ptr, d_toa%p)
    if (rc == 0) rc = d_toa%download(c_loc(toa_src), int(size(toa_src), c_size_t))
    error_msg = rrnn_error_msg(rc)
    call d_play%free(); call d_plev%free(); call d_tlay%free(); call d_toa%free()
    if (allocated(gas_bufs)) then
      do i = 1, size(gas_bufs)
        call gas_bufs(i)%free()
      end do
    end if
  end function gas_optics_ext

end module mo_gas_optics_rrtmgp

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_rte_lw.F90:60-424
module mo_rte_lw
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  use mo_optical_props, only: ty_optical_props_arry, ty_optical_props_1scl, ty_optical_props_2str
  use mo_source_functions, only: ty_source_func_lw
  use mo_fluxes, only: ty_fluxes_flexible
  implicit none
  private
  public :: rte_lw

contains

  function rte_lw(optical_props, top_at_1, &
                  sources, sfc_emis,       &
                  fluxes,                  &
                  inc_flux, n_gauss_angles, use_2stream, &
                  lw_Ds, flux_up_Jac, flux_dn_Jac) result(error_msg)
    class(ty_optical_props_arry), intent(in   ) :: optical_props
    logical,                      intent(in   ) :: top_at_1
    type(ty_source_func_lw),      intent(in   ) :: sources
    real(wp), dimension(:,:), contiguous, target, intent(in) :: sfc_emis          ! (nband, ncol)
    class(ty_fluxes_flexible),    intent(inout) :: fluxes
    real(wp), dimension(:,:), contiguous, target, optional, intent(in   ) :: inc_flux         ! (ngpt, ncol)
    integer,                  optional, intent(in   ) :: n_gauss_angles
    logical,                  optional, intent(in   ) :: use_2stream
    real(wp), dimension(:,:), contiguous, target, optional, intent(in   ) :: lw_Ds            ! (ngpt, ncol)
    real(wp), dimension(:,:), contiguous, target, optional, intent(inout) :: flux_up_Jac, flux_dn_Jac
    character(len=128) :: error_msg
    integer :: ncol, nlay, ngpt, nband, n_quad_angs
    integer(c_int) :: rc, t1
    integer(c_size_t) :: nlev_all
    logical :: general, two_stream
    type(ty_devbuf) :: d_emis, d_inc, d_Ds, d_up, d_dn, d_jac, d_gup, d_gdn, d_net
    type(c_ptr) :: p_inc, p_Ds, p_ssa, p_g, p_jac, p_gup, p_gdn

    ncol  = optical_props%get_ncol()
    nlay  = optical_props%get_nlay()
    ngpt  = optical_props%get_ngpt()
    nband = optical_props%get_nband()
    nlev_all = int(nlay + 1, c_size_t) * ncol
    error_msg = ""
    if (.not. fluxes%are_desired()) error_msg = "rte_lw: no space allocated for fluxes"
    if (sources%get_ncol() /= ncol .or. sources%get_nlay() /= nlay .or. sources%get_ngpt() /= ngpt) &
      error_msg = "rte_lw: sources and optical properties inconsistently sized"
    if (size(sfc_emis, 1) /= nband .or. size(sfc_emis, 2) /= ncol) error_msg = "rte_lw: sfc_emis inconsistently sized"
    if (any(sfc_emis < 0._wp .or. sfc_emis > 1._wp)) error_msg = "rte_lw: sfc_emis has values < 0 or > 1"
    if (present(inc_flux)) then
      if (size(inc_flux, 1) /= ngpt .or. size(inc_flux, 2) /= ncol) error_msg = "rte_lw: inc_flux inconsistently sized"
    end if
    n_quad_angs = 1
    if (present(n_gauss_angles)) then
      if (n_gauss_angles > 4) error_msg = "rte_lw: asking for too many quadrature points for no-scattering calculation"
      if (n_gauss_angles < 1) error_msg = "rte_lw: have to ask for at least one quadrature point for no-scattering calculation"
      n_quad_angs = n_gauss_angles
    end if
    if (present(lw_Ds)) then
      if (size(lw_Ds, 1) /= ngpt .or. size(lw_Ds, 2) /= ncol) error_msg = "rte_lw: lw_Ds inconsistently sized"
      if (any(lw_Ds < 1._wp)) error_msg = "rte_lw: one or more values of lw_Ds < 1."
      if (n_quad_angs /= 1) error_msg = "rte_lw: providing lw_Ds incompatible with specifying n_gauss_angles"
    end if
    two_stream = .false.
    if (present(use_2stream)) two_stream = use_2stream
    if (present(flux_dn_Jac)) error_msg = "rte_lw: flux_dn_Jac is not computed by this fork's solvers"
    if (error_msg /= "") return

    t1 = merge(1_c_int, 0_c_int, top_at_1)
    rc = d_emis%upload(c_loc(sfc_emis), int(size(sfc_emis), c_size_t))
    p_inc = c_null_ptr
    if (rc == 0 .and. present(inc_flux)) then
      rc = d_inc%upload(c_loc(inc_flux), int(size(inc_flux), c_size_t))
      p_inc = d_inc%p
    end if
    p_Ds = c_null_ptr
    if (rc == 0 .and. present(lw_Ds)) then
      rc = d_Ds%upload(c_loc(lw_Ds), int(size(lw_Ds), c_size_t))
      p_Ds = d_Ds%p
    end if
    if (rc == 0) rc = d_up%resize(nlev_all)
    if (rc == 0) rc = d_dn%resize(nlev_all)
    p_jac = c_null_ptr
    if (rc == 0 .and. present(flux_up_Jac)) then
      rc = d_jac%resize(nlev_all)
      p_jac = d_jac%p
    end if
    p_gup = c_null_ptr
    p_gdn = c_null_ptr
    if (rc == 0 .and. fluxes%are_desired_gpt()) then
      rc = d_gup%resize(nlev_all * ngpt)
      if (rc == 0) rc = d_gdn%resize(nlev_all * ngpt)
      p_gup = d_gup%p
      p_gdn = d_gdn%p
    end if
    p_ssa = c_null_ptr
    p_g = c_null_ptr
    select type (optical_props)
    class is (ty_optical_props_2str)
      p_ssa = optical_props%ssa%p
      p_g = optical_props%g%p
    end select
    general = present(lw_Ds) .or. present(flux_up_Jac) .or. fluxes%are_desired_gpt() .or. c_associated(p_ssa)

    if (rc == 0) then
      if (two_stream .and. c_associated(p_ssa)) then          ! lw_solver_2stream, rte/mo_rte_lw.F90:322-352
        rc = rrnn_rte_lw_2stream(rrnn_ctx(), optical_props%kd, int(nlay, c_int), int(ncol, c_int), t1, p_inc, &
                                 optical_props%tau%p, p_ssa, p_g, sources%lev_source%p, sources%sfc_source%p, d_emis%p, &
                                 d_up%p, d_dn%p, p_gup, p_gdn)
      else if (general) then                                  ! re-scaled transport / lw_Ds / Jacobian / g-point fluxes (:278-320)
        rc = rrnn_rte_lw_ext(rrnn_ctx(), optical_props%kd, int(nlay, c_int), int(ncol, c_int), t1, int(n_quad_angs, c_int), &
                             p_inc, optical_props%tau%p, p_ssa, p_g, sources%lay_source%p, sources%lev_source%p, &
                             sources%sfc_source%p, d_emis%p, p_Ds, sources%sfc_source_Jac%p, d_up%p, d_dn%p, p_jac, p_gup, p_gdn)
      else                                                    ! lw_solver_noscat_GaussQuad: the tuned kernel
        rc = rrnn_rte_lw(rrnn_ctx(), optical_props%kd, int(nlay, c_int), int(ncol, c_int), t1, int(n_quad_angs, c_int), p_inc, &
                         optical_props%tau%p, sources%lay_source%p, sources%lev_source%p, sources%sfc_source%p, d_emis%p, &
                         d_up%p, d_dn%p)
      end if
    end if
    ! ty_fluxes_broadband%reduce (rte/mo_fluxes.F90:97-170): whatever the caller associated
    if (rc == 0 .and. associated(fluxes%flux_up)) rc = d_up%download(c_loc(fluxes%flux_up), nlev_all)
    if (rc == 0 .and. associated(fluxes%flux_dn)) rc = d_dn%download(c_loc(fluxes%flux_dn), nlev_all)
    if (rc == 0 .and. associated(fluxes%flux_net)) then
      rc = d_net%resize(nlev_all)
      if (rc == 0) rc = rrnn_net_flux(rrnn_ctx(), nlev_all, d_dn%p, d_up%p, d_net%p)
      if (rc == 0) rc = d_net%download(c_loc(fluxes%flux_net), nlev_all)
    end if
    if (rc == 0 .and. present(flux_up_Jac)) rc = d_jac%download(c_loc(flux_up_Jac), nlev_all)
    if (rc == 0 .and. associated(fluxes%gpt_flux_up)) rc = d_gup%download(c_loc(fluxes%gpt_flux_up), nlev_all * ngpt)
    if (rc == 0 .and. associated(fluxes%gpt_flux_dn)) rc = d_gdn%download(c_loc(fluxes%gpt_flux_dn), nlev_all * ngpt)
    error_msg = rrnn_error_msg(rc)
    call d_emis%free(); call d_inc%free(); call d_Ds%free(); call d_up%free(); call d_dn%free()
    call d_jac%free(); call d_gup%free(); call d_gdn%free(); call d_net%free()
  end function rte_lw

end module mo_rte_lw

! ------------------------------------------------------------------------------------------------------------------------------
! rte/mo_rte_sw.F90:48-266
module mo_rte_sw
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  use mo_optical_props, only: ty_optical_props_arry, ty_optical_props_2str
  use mo_fluxes, only: ty_fluxes_flexible
  implicit none
  private
  public :: rte_sw

contains

  function rte_sw(atmos, top_at_1,                 &
                  mu0, inc_flux,                   &
                  sfc_alb_dir_gpt, sfc_alb_dif_gpt,        &
                  fluxes, inc_flux_dif &
                  ) result(error_msg)
    class(ty_optical_props_arry), intent(in   ) :: atmos
    logical,                      intent(in   ) :: top_at_1
    real(wp), dimension(:),   contiguous, target, intent(in) :: mu0                              ! (ncol)
    real(wp), dimension(:,:), contiguous, target, intent(in) :: inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt   ! (ngpt, ncol)
    class(ty_fluxes_flexible),    intent(inout) :: fluxes
    real(wp), dimension(:,:), optional, contiguous, target, intent(in) :: inc_flux_dif
    character(len=128) :: error_msg
    integer :: ncol, nlay, ngpt
    integer(c_int) :: rc, t1
    integer(c_size_t) :: nlev_all
    type(ty_devbuf) :: d_mu0, d_inc, d_adir, d_adif, d_incdif, d_up, d_dn, d_dir, d_net, d_gup, d_gdn, d_gdir
    type(c_ptr) :: p_incdif, p_g

    ncol = atmos%get_ncol()
    nlay = atmos%get_nlay()
    ngpt = atmos%get_ngpt()
    nlev_all = int(nlay + 1, c_size_t) * ncol
    error_msg = ""
    if (.not. fluxes%are_desired()) error_msg = "rte_sw: no space allocated for fluxes"
    if (size(mu0) /= ncol) error_msg = "rte_sw: mu0 inconsistently sized"
    if (any(mu0 < 0._wp .or. mu0 > 1._wp)) error_msg = "rte_sw: one or more mu0 <= 0 or > 1"
    if (size(inc_flux, 1) /= ngpt .or. size(inc_flux, 2) /= ncol) error_msg = "rte_sw: inc_flux inconsistently sized"
    if (any(inc_flux < 0._wp)) error_msg = "rte_sw: one or more inc_flux < 0"
    if (size(sfc_alb_dir_gpt, 1) /= ngpt .or. size(sfc_alb_dir_gpt, 2) /= ncol) error_msg = "rte_sw: sfc_alb_dir inconsistently sized"
    if (size(sfc_alb_dif_gpt, 1) /= ngpt .or. size(sfc_alb_dif_gpt, 2) /= ncol) error_msg = "rte_sw: sfc_alb_dif inconsistently sized"
    if (any(sfc_alb_dir_gpt < 0._wp .or. sfc_alb_dir_gpt > 1._wp)) error_msg = "rte_sw: sfc_alb_dir out of bounds [0,1]"
    if (any(sfc_alb_dif_gpt < 0._wp .or. sfc_alb_dif_gpt > 1._wp)) error_msg = "rte_sw: sfc_alb_dif out of bounds [0,1]"
    if (present(inc_flux_dif)) then
      if (size(inc_flux_dif, 1) /= ngpt .or. size(inc_flux_dif, 2) /= ncol) error_msg = "rte_sw: inc_flux_dif inconsistently sized"
      if (any(inc_flux_dif < 0._wp)) error_msg = "rte_sw: one or more inc_flux_dif < 0"
    end if
    if (error_msg /= "") return

    select type (atmos)
    class is (ty_optical_props_2str)
      t1 = merge(1_c_int, 0_c_int, top_at_1)
      rc = d_mu0%upload(c_loc(mu0), int(size(mu0), c_size_t))
      if (rc == 0) rc = d_inc%upload(c_loc(inc_flux), int(size(inc_flux), c_size_t))
      if (rc == 0) rc = d_adir%upload(c_loc(sfc_alb_dir_gpt), int(size(sfc_alb_dir_gpt), c_size_t))
      if (rc == 0) rc = d_adif%upload(c_loc(sfc_alb_dif_gpt), int(size(sfc_alb_dif_gpt), c_size_t))
      p_incdif = c_null_ptr
      if (rc == 0 .and. present(inc_flux_dif)) then
        rc = d_incdif%upload(c_loc(inc_flux_dif), int(size(inc_flux_dif), c_size_t))
        p_incdif = d_incdif%p
      end if
      if (rc == 0) rc = d_up%resize(nlev_all)
      if (rc == 0) rc = d_dn%resize(nlev_all)
      if (rc == 0) rc = d_dir%resize(nlev_all)
      p_g = atmos%g%p                                           ! NULL on the clear-sky NN path: g == 0 stays implicit
      if (rc == 0) then
        if (fluxes%are_desired_gpt()) then                      ! sw_solver_2stream with save_gpt_flux (:541-692)
          rc = d_gup%resize(nlev_all * ngpt)
          if (rc == 0) rc = d_gdn%resize(nlev_all * ngpt)
          if (rc == 0) rc = d_gdir%resize(nlev_all * ngpt)
          if (rc == 0) rc = rrnn_sw_solver_2stream_ext(rrnn_ctx(), int(ngpt, c_int), int(nlay, c_int), int(ncol, c_int), t1, &
                                 d_inc%p, p_incdif, atmos%tau%p, atmos%ssa%p, p_g, d_mu0%p, d_adir%p, d_adif%p, &
                                 d_up%p, d_dn%p, d_dir%p, d_gup%p, d_gdn%p, d_gdir%p)
        else
          rc = rrnn_rte_sw(rrnn_ctx(), int(ngpt, c_int), int(nlay, c_int), int(ncol, c_int), t1, d_mu0%p, d_inc%p, &
                           d_adir%p, d_adif%p, p_incdif, atmos%tau%p, atmos%ssa%p, p_g, d_up%p, d_dn%p, d_dir%p)
        end if
      end if
      if (rc == 0 .and. associated(fluxes%flux_up)) rc = d_up%download(c_loc(fluxes%flux_up), nlev_all)
      if (rc == 0 .and. associated(fluxes%flux_dn)) rc = d_dn%download(c_loc(fluxes%flux_dn), nlev_all)
      if (rc == 0 .and. associated(fluxes%flux_dn_dir)) rc = d_dir%download(c_loc(fluxes%flux_dn_dir), nlev_all)
      if (rc == 0 .and. associated(fluxes%flux_net)) then
        rc = d_net%resize(nlev_all)
        if (rc == 0) rc = rrnn_net_flux(rrnn_ctx(), nlev_all, d_dn%p, d_up%p, d_net%p)
        if (rc == 0) rc = d_net%download(c_loc(fluxes%flux_net), nlev_all)
      end if
      if (rc == 0 .and. associated(fluxes%gpt_flux_up)) rc = d_gup%download(c_loc(fluxes%gpt_flux_up), nlev_all * ngpt)
      if (rc == 0 .and. associated(fluxes%gpt_flux_dn)) rc = d_gdn%download(c_loc(fluxes%gpt_flux_dn), nlev_all * ngpt)
      if (rc == 0 .and. associated(fluxes%gpt_flux_dn_dir)) rc = d_gdir%download(c_loc(fluxes%gpt_flux_dn_dir), nlev_all * ngpt)
      error_msg = rrnn_error_msg(rc)
      call d_mu0%free(); call d_inc%free(); call d_adir%free(); call d_adif%free(); call d_incdif%free()
      call d_up%free(); call d_dn%free(); call d_dir%free(); call d_net%free()
      call d_gup%free(); call d_gdn%free(); call d_gdir%free()
    class default
      ! rte/mo_rte_sw.F90:207-215: no solar source function is coded for the no-scattering case
      error_msg = "rte_sw: shortwave calculations require two-stream optical properties (ty_optical_props_2str)"
    end select
  end function rte_sw

end module mo_rte_sw

! ------------------------------------------------------------------------------------------------------------------------------
! extensions/mo_heating_rates.F90:26-54 (this fork's (nlay+1, ncol) layout)
module mo_heating_rates
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  implicit none
  private
  public :: compute_heating_rate

contains

  function compute_heating_rate(flux_up, flux_dn, plev, heating_rate) result(error_msg)
    real(wp), dimension(:,:), contiguous, target, intent(in ) :: flux_up, flux_dn, plev      ! (nlay+1, ncol)
    real(wp), dimension(:,:), contiguous, target, intent(out) :: heating_rate                ! (nlay, ncol)
    character(len=128) :: error_msg
    integer :: ncol, nlay
    integer(c_int) :: rc
    type(ty_devbuf) :: d_up, d_dn, d_p, d_hr
    error_msg = ""
    nlay = size(flux_up, 1) - 1
    ncol = size(flux_up, 2)
    if (size(flux_dn, 1) /= nlay + 1 .or. size(flux_dn, 2) /= ncol) error_msg = "heating_rate: flux_dn array inconsistently sized."
    if (size(plev, 1) /= nlay + 1 .or. size(plev, 2) /= ncol) error_msg = "heating_rate: plev array inconsistently sized."
    if (size(heating_rate, 1) /= nlay .or. size(heating_rate, 2) /= ncol) &
      error_msg = "heating_rate: heating_rate array inconsistently sized."
    if (error_msg /= "") return
    rc = d_up%upload(c_loc(flux_up), int(size(flux_up), c_size_t))
    if (rc == 0) rc = d_dn%upload(c_loc(flux_dn), int(size(flux_dn), c_size_t))
    if (rc == 0) rc = d_p%upload(c_loc(plev), int(size(plev), c_size_t))
    if (rc == 0) rc = d_hr%resize(int(size(heating_rate), c_size_t))
    if (rc == 0) rc = rrnn_heating_rate(rrnn_ctx(), int(ncol, c_int), int(nlay, c_int), d_up%p, d_dn%p, d_p%p, d_hr%p)
    if (rc == 0) rc = d_hr%download(c_loc(heating_rate), int(size(heating_rate), c_size_t))
    error_msg = rrnn_error_msg(rc)
    call d_up%free(); call d_dn%free(); call d_p%free(); call d_hr%free()
  end function compute_heating_rate

end module mo_heating_rates

! ------------------------------------------------------------------------------------------------------------------------------
! extensions/cloud_optics/mo_cloud_optics.F90:32-170, 354-535 (LUT branch)
module mo_cloud_optics
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  use mo_rrnn_device, only: rrnn_ctx, ty_devbuf
  use mo_optical_props, only: ty_optical_props, ty_optical_props_arry, ty_optical_props_1scl, ty_optical_props_2str
  implicit none
  private
  public :: ty_cloud_optics

  type, extends(ty_optical_props) :: ty_cloud_optics
    type(c_ptr) :: lut = c_null_ptr
    integer     :: icergh = 0
    real(wp)    :: radliq_lwr = 0._wp, radliq_upr = 0._wp, radice_lwr = 0._wp, radice_upr = 0._wp
    ! the tables as loaded, (nsize, nbnd, nrghice) for ice, until set_ice_roughness picks one (mo_cloud_optics.F90:323-349)
    real(wp), allocatable :: extliq(:,:), ssaliq(:,:), asyliq(:,:), extice(:,:,:), ssaice(:,:,:), asyice(:,:,:)
  contains
    procedure, public :: load_lut
    procedure, public :: set_ice_roughness
    procedure, public :: cloud_optics
    procedure, public :: get_min_radius_liq, get_max_radius_liq, get_min_radius_ice, get_max_radius_ice
    procedure, public :: finalize => cloud_optics_finalize
  end type ty_cloud_optics

contains

  function load_lut(this, band_lims_wvn, radliq_lwr, radliq_upr, radliq_fac, radice_lwr, radice_upr, radice_fac, &
                    lut_extliq, lut_ssaliq, lut_asyliq, lut_extice, lut_ssaice, lut_asyice) result(error_msg)
    class(ty_cloud_optics),     intent(inout) :: this
    real(wp), dimension(:,:),   intent(in   ) :: band_lims_wvn
    real(wp),                   intent(in   ) :: radliq_lwr, radliq_upr, radliq_fac, radice_lwr, radice_upr, radice_fac
    real(wp), dimension(:,:),   intent(in   ) :: lut_extliq, lut_ssaliq, lut_asyliq        ! (nsize_liq, nbnd)
    real(wp), dimension(:,:,:), intent(in   ) :: lut_extice, lut_ssaice, lut_asyice        ! (nsize_ice, nbnd, nrghice)
    character(len=128) :: error_msg
    error_msg = this%init(band_lims_wvn, name="RRTMGP cloud optics")
    if (error_msg /= "") return
    if (size(lut_extliq, 2) /= size(band_lims_wvn, 2)) then
      error_msg = "cloud_optics%init(): arrays have inconsistent sizes"
      return
    end if
    this%radliq_lwr = radliq_lwr; this%radliq_upr = radliq_upr
    this%radice_lwr = radice_lwr; this%radice_upr = radice_upr
    this%extliq = lut_extliq; this%ssaliq = lut_ssaliq; this%asyliq = lut_asyliq
    this%extice = lut_extice; this%ssaice = lut_ssaice; this%asyice = lut_asyice
    error_msg = this%set_ice_roughness(1)
  end function load_lut

  function set_ice_roughness(this, icergh) result(error_msg)
    class(ty_cloud_optics), target, intent(inout) :: this
    integer,                intent(in   ) :: icergh
    character(len=128) :: error_msg
    integer(c_int) :: rc
    real(wp), allocatable, target :: e(:,:), s(:,:), a(:,:)
    error_msg = ""
    if (.not. allocated(this%extice)) then
      error_msg = "cloud_optics%set_ice_roughness(): can't set before initialization"
    else if (icergh < 1 .or. icergh > size(this%extice, 3)) then
      error_msg = "cloud optics: cloud ice surface roughness flag is out of bounds"
    end if
    if (error_msg /= "") return
    this%icergh = icergh
    if (c_associated(this%lut)) rc = rrnn_cloud_lut_destroy(this%lut)
    e = this%extice(:, :, icergh); s = this%ssaice(:, :, icergh); a = this%asyice(:, :, icergh)
    rc = rrnn_cloud_lut_create(rrnn_ctx(), int(size(this%extliq, 2), c_int), int(size(this%extliq, 1), c_int), &
                               int(size(e, 1), c_int), this%radliq_lwr, this%radliq_upr, this%radice_lwr, this%radice_upr, &
                               c_loc(this%extliq), c_loc(this%ssaliq), c_loc(this%asyliq), c_loc(e), c_loc(s), c_loc(a), this%lut)
    error_msg = rrnn_error_msg(rc)
  end function set_ice_roughness

  ! mo_cloud_optics.F90:354-535: by-band optical properties of the clouds; masks are where the water paths are > 0
  function cloud_optics(this, clwp, ciwp, reliq, reice, optical_props) result(error_msg)
    class(ty_cloud_optics), intent(in   ) :: this
    real(wp), dimension(:,:), contiguous, target, intent(in) :: clwp, ciwp, reliq, reice     ! (nlay, ncol) in this fork
    class(ty_optical_props_arry), intent(inout) :: optical_props
    character(len=128) :: error_msg
    integer :: ncol, nlay
    integer(c_int) :: rc
    type(ty_devbuf) :: d_clwp, d_ciwp, d_reliq, d_reice
    type(c_ptr) :: p_ssa, p_g
    error_msg = ""
    nlay = size(clwp, 1)
    ncol = size(clwp, 2)
    if (.not. c_associated(this%lut)) error_msg = "cloud optics: no data has been initialized"
    if (optical_props%get_ncol() /= ncol .or. optical_props%get_nlay() /= nlay) &
      error_msg = "cloud optics: optical properties have the wrong extents"
    if (optical_props%get_ngpt() /= this%get_nband()) error_msg = "cloud optics: optical properties must be defined by band"
    if (any(shape(ciwp) /= shape(clwp)) .or. any(shape(reliq) /= shape(clwp)) .or. any(shape(reice) /= shape(clwp))) &
      error_msg = "cloud optics: ciwp, reliq or reice has wrong extents"
    if (any(clwp > 0._wp .and. (reliq < this%radliq_lwr .or. reliq > this%radliq_upr))) &
      error_msg = "cloud optics: liquid effective radius is out of bounds"
    if (any(ciwp > 0._wp .and. (reice < this%radice_lwr .or. reice > this%radice_upr))) &
      error_msg = "cloud optics: ice effective radius is out of bounds"
    if (any(clwp < 0._wp) .or. any(ciwp < 0._wp)) error_msg = "cloud optics: negative clwp or ciwp where clouds are supposed to be"
    if (error_msg /= "") return
    rc = d_clwp%upload(c_loc(clwp), int(size(clwp), c_size_t))
    if (rc == 0) rc = d_ciwp%upload(c_loc(ciwp), int(size(ciwp), c_size_t))
    if (rc == 0) rc = d_reliq%upload(c_loc(reliq), int(size(reliq), c_size_t))
    if (rc == 0) rc = d_reice%upload(c_loc(reice), int(size(reice), c_size_t))
    p_ssa = c_null_ptr
    p_g = c_null_ptr
    select type (optical_props)
    class is (ty_optical_props_2str)
      if (rc == 0 .and. .not. optical_props%g%is_alloc()) rc = optical_props%g%resize(optical_props%tau%n)
      p_ssa = optical_props%ssa%p
      p_g = optical_props%g%p
    end select
    if (rc == 0) rc = rrnn_cloud_optics(rrnn_ctx(), this%lut, int(ncol, c_int), int(nlay, c_int), d_clwp%p, d_ciwp%p, d_reliq%p, &
                                        d_reice%p, optical_props%tau%p, p_ssa, p_g)
    error_msg = rrnn_error_msg(rc)
    call d_clwp%free(); call d_ciwp%free(); call d_reliq%free(); call d_reice%free()
  end function cloud_optics

  function get_min_radius_liq(this) result(r)
    class(ty_cloud_optics), intent(in) :: this
    real(wp) :: r
    r = this%radliq_lwr
  end function get_min_radius_liq
  function get_max_radius_liq(this) result(r)
    class(ty_cloud_optics), intent(in) :: this
    real(wp) :: r
    r = this%radliq_upr
  end function get_max_radius_liq
  function get_min_radius_ice(this) result(r)
    class(ty_cloud_optics), intent(in) :: this
    real(wp) :: r
    r = this%radice_lwr
  end function get_min_radius_ice
  function get_max_radius_ice(this) result(r)
    class(ty_cloud_optics), intent(in) :: this
    real(wp) :: r
    r = this%radice_upr
  end function get_max_radius_ice

  subroutine cloud_optics_finalize(this)
    class(ty_cloud_optics), intent(inout) :: this
    integer(c_int) :: rc
    if (c_associated(this%lut)) rc = rrnn_cloud_lut_destroy(this%lut)
    this%lut = c_null_ptr
    if (allocated(this%extliq)) deallocate(this%extliq, this%ssaliq, this%asyliq, this%extice, this%ssaice, this%asyice)
    call this%finalize_base()
  end subroutine cloud_optics_finalize

end module mo_cloud_optics

! =====================================================================================================================
! extensions/solar_variability/mo_solar_variability.F90:20-183 -- ty_solar_var: the mean-solar-cycle table of the facular
! (mg) and sunspot (sb) indices and its interpolation to a cycle fraction; the pair feeds set_solar_variability.
module mo_solar_variability
  use, intrinsic :: iso_c_binding
  use mo_rte_kind, only: wp
  use mo_rrnn_c_binding
  implicit none
  private

  type, public :: ty_solar_var
    real(wp), dimension(:,:), allocatable :: avgcyc_ind      ! (nsolarterms, nsolarfrac) -> (2,134)
  contains
    procedure, public :: solar_var_ind_interp
    procedure, public :: load
    procedure, public :: finalize
  end type ty_solar_var

contains

  function load(this, avgcyc_ind) result(error_msg)      ! :45-69
    class(ty_solar_var),      intent(inout) :: this
    real(wp), dimension(:,:), intent(in   ) :: avgcyc_ind
    character(len=128)    :: error_msg
    error_msg = ""
    if (allocated(this%avgcyc_ind)) deallocate(this%avgcyc_ind)
    allocate(this%avgcyc_ind(size(avgcyc_ind, dim=1), size(avgcyc_ind, dim=2)))
    this%avgcyc_ind = avgcyc_ind
  end function load

  subroutine finalize(this)      ! :75-83
    class(ty_solar_var), intent(inout) :: this
    if (allocated(this%avgcyc_ind)) deallocate(this%avgcyc_ind)
  end subroutine finalize

  function solar_var_ind_interp(this, solcycfrac, mg_index, sb_index) result(error_msg)      ! :91-183
    class(ty_solar_var), target, intent(in   ) :: this
    real(wp),            intent(in   ) :: solcycfrac
    real(wp), target,    intent(out  ) :: mg_index
    real(wp), target,    intent(out  ) :: sb_index
    character(len=128)                 :: error_msg
    error_msg = ""
    if (solcycfrac .lt. 0._wp .or. solcycfrac .gt. 1._wp) error_msg = 'solar_var_ind_interp: solcycfrac out of range'
    if (error_msg /= '') return
    if (allocated(this%avgcyc_ind)) then      ! (2, nsolarfrac) in memory == C [nsolarfrac][2]
      error_msg = rrnn_error_msg(rrnn_solar_var_ind_interp(c_loc(this%avgcyc_ind), int(size(this%avgcyc_ind, 2), c_int), solcycfrac, &
                                                           c_loc(mg_index), c_loc(sb_index)))
    end if
  end function solar_var_ind_interp

end module mo_solar_variability
