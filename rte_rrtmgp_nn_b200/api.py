"""Host-side mirror of the reference's Fortran API for the NN gas optics + RTE path.

Same names, argument meaning and error behaviour as the reference (functions return the reference's
`error_msg` string, empty on success), bodies forward to the C ABI (include/rrnn.h).  Device arrays are
torch CUDA tensors (PyTorch is only the allocator / stream provider here); layouts are the reference's with
the g-point fastest: tau[ncol, nlay, ngpt] == Fortran tau(ngpt, nlay, ncol).

  rrtmgp_network_type      neural/mod_network_rrtmgp.F90:34-53
  ty_gas_concs             rrtmgp/mo_gas_concentrations.F90:50-88
  ty_gas_optics_rrtmgp     rrtmgp/mo_gas_optics_rrtmgp.F90:61-200 (gas_optics :239-243, :433-437)
  ty_optical_props_1scl/_2str  rte/mo_optical_props.F90:98-192
  ty_source_func_lw        rte/mo_source_functions.F90:26-43
  ty_fluxes_broadband      rte/mo_fluxes.F90:46-67
  rte_lw / rte_sw          rte/mo_rte_lw.F90:60-64, rte/mo_rte_sw.F90:48-52
  ty_cloud_optics          extensions/cloud_optics/mo_cloud_optics.F90:32-70, :354-535
  compute_heating_rate     extensions/mo_heating_rates.F90:26-54
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import RRNNError, rrnn_gas_t, vp

ACT_NAMES = ["linear", "softsign", "relu", "sigmoid", "hard_sigmoid"]


def _torch():
    import torch
    return torch


class Context:
    """rrnn_ctx_t: one per GPU / host thread; calls on one context are stream-ordered."""

    def __init__(self, device=0, stream=None):
        self.h = vp()
        _lib.check(_lib.lib().rrnn_ctx_create(int(device), vp(stream) if stream else None, C.byref(self.h)))
        self.device = int(device)

    def synchronize(self):
        _lib.check(_lib.lib().rrnn_ctx_synchronize(self.h))

    def set_flag(self, name, value):
        _lib.check(_lib.lib().rrnn_ctx_set_flag(self.h, name.encode(), int(value)))

    def set_chunk_columns(self, n):
        _lib.check(_lib.lib().rrnn_ctx_set_chunk_columns(self.h, int(n)))

    KERNEL_KINDS = ("gas_optics_lw", "lw_solver", "gas_optics_sw", "sw_solver")

    def profile(self, enable=True):
        _lib.check(_lib.lib().rrnn_ctx_profile(self.h, int(enable)))

    def profile_read(self):
        """{kernel: (total_ms, launches)} since profile(True)."""
        out = {}
        for k, name in enumerate(self.KERNEL_KINDS):
            ms = C.c_double(0); n = C.c_int(0)
            _lib.check(_lib.lib().rrnn_ctx_profile_read(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    @property
    def stream(self):
        return _lib.lib().rrnn_ctx_stream(self.h)

    @property
    def launch_count(self):
        return int(_lib.lib().rrnn_ctx_launch_count(self.h))

    NN_KERNELS = {0: "none", 1: "ffma", 2: "tcgen05"}

    @property
    def last_nn_kernel(self):
        """Which MLP kernel served the most recent NN gas-optics call: "none", "ffma" or "tcgen05"."""
        return self.NN_KERNELS[int(_lib.lib().rrnn_ctx_last_nn_kernel(self.h))]

    @property
    def nn_kernel_counts(self):
        """(tcgen05 launches, fp32 FFMA launches) of the NN gas optics since the context was created."""
        a = C.c_longlong(0); b = C.c_longlong(0)
        _lib.check(_lib.lib().rrnn_ctx_nn_kernel_counts(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def torch_stream(self):
        torch = _torch()
        if not self.stream:  # legacy default stream == torch's default stream
            return torch.cuda.default_stream(torch.device("cuda", self.device))
        return torch.cuda.ExternalStream(self.stream, device=torch.device("cuda", self.device))

    def close(self):
        if self.h:
            _lib.lib().rrnn_ctx_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _dev(t, ctx=None):
    """float32 contiguous CUDA tensor from tensor / ndarray / None."""
    if t is None:
        return None
    torch = _torch()
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    dev = torch.device("cuda", ctx.device if ctx else 0)
    if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=dev, dtype=torch.float32).contiguous()
    return t


def _ptr(t):
    return vp(t.data_ptr()) if t is not None else None


# rte/mo_rte_rrtmgp_config.F90:23-24, 52-67: run-time checks, off by default, switched by rte_rrtmgp_config_checks
check_extents = False
check_values = False


def rte_rrtmgp_config_checks(extents, values=None):
    """rte_rrtmgp_config_checks(extents, values) / rte_rrtmgp_config_checks(do_checks) (mo_rte_rrtmgp_config.F90:52-67)."""
    global check_extents, check_values
    check_extents = bool(extents)
    check_values = bool(extents if values is None else values)


def _outside(t, lo, hi):
    return bool((t < lo).any() or (t > hi).any())


class rrtmgp_network_type:
    """The reference's network container; `load_netcdf` reads the shipped netCDF-4 weight files."""

    def __init__(self, ctx=None):
        self.ctx = ctx
        self.h = vp()

    def load_netcdf(self, filename):
        ctxh = self.ctx.h if self.ctx is not None else None
        _lib.check(_lib.lib().rrnn_model_load_netcdf(ctxh, str(filename).encode(), C.byref(self.h)))
        return self

    def load(self, model_txt, scaling_txt):
        ctxh = self.ctx.h if self.ctx is not None else None
        _lib.check(_lib.lib().rrnn_model_load_ascii(ctxh, str(model_txt).encode(), str(scaling_txt).encode(), C.byref(self.h)))
        return self

    def save(self, model_txt, scaling_txt):
        _lib.check(_lib.lib().rrnn_model_save_ascii(self.h, str(model_txt).encode(), str(scaling_txt).encode()))

    @classmethod
    def from_arrays(cls, ctx, model):
        """model: dict as produced by oracle.nc4min.load_nn_model (dims, W, b, activations, ...)."""
        self = cls(ctx)
        dims = np.asarray(model["dims"], np.int32)
        wpack = np.concatenate([np.ascontiguousarray(w, np.float32).ravel() for w in model["W"]])
        bpack = np.concatenate([np.asarray(b, np.float32).ravel() for b in model["b"]])
        act = np.asarray([ACT_NAMES.index(a) for a in model["activations"]], np.int32)
        names = b"".join(n.encode().ljust(32) for n in model["input_names"])
        fp = lambda a: None if a is None else np.ascontiguousarray(a, np.float32).ctypes.data_as(_lib.c_float_p)
        keep = [np.ascontiguousarray(model[k], np.float32) if model.get(k) is not None else None
                for k in ("xmin", "xmax", "ymean", "ystd")]
        _lib.check(_lib.lib().rrnn_model_create(ctx.h if ctx else None, len(dims) - 1, dims.ctypes.data_as(_lib.c_int_p),
                                                wpack.ctypes.data_as(_lib.c_float_p), bpack.ctypes.data_as(_lib.c_float_p),
                                                act.ctypes.data_as(_lib.c_int_p), fp(keep[0]), fp(keep[1]), fp(keep[2]),
                                                fp(keep[3]), names, C.byref(self.h)))
        return self

    @property
    def dims(self):
        n = _lib.lib().rrnn_model_nlayers(self.h)
        d = (C.c_int * (n + 1))()
        _lib.check(_lib.lib().rrnn_model_dims(self.h, d))
        return list(d)

    @property
    def input_names(self):
        out = []
        for i in range(self.dims[0]):
            b = C.create_string_buffer(32)
            _lib.check(_lib.lib().rrnn_model_input_name(self.h, i, b))
            out.append(b.value.decode())
        return out

    @property
    def activations(self):
        return [ACT_NAMES[_lib.lib().rrnn_model_activation(self.h, l)] for l in range(len(self.dims) - 1)]

    def _get(self, which, layer=0):
        n = C.c_int(0)
        _lib.check(_lib.lib().rrnn_model_get(self.h, which, layer, None, C.byref(n)))
        if n.value == 0:
            return None
        a = np.empty(n.value, np.float32)
        _lib.check(_lib.lib().rrnn_model_get(self.h, which, layer, a.ctypes.data_as(_lib.c_float_p), C.byref(n)))
        return a

    def weights(self, layer):
        d = self.dims
        return self._get(0, layer).reshape(d[layer], d[layer + 1])

    def bias(self, layer):
        return self._get(1, layer)

    coeffs_input_min = property(lambda self: self._get(2))
    coeffs_input_max = property(lambda self: self._get(3))
    coeffs_output_mean = property(lambda self: self._get(4))
    coeffs_output_std = property(lambda self: self._get(5))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rrnn_model_destroy(self.h)
        except Exception:
            pass


class ty_gas_concs:
    """Gas volume mixing ratios by name: scalar, per-layer profile (nlay) or full field (ncol, nlay)."""

    def __init__(self, gas_names=None):
        self.gas_name = []
        self.concs = {}
        if gas_names:
            self.init(gas_names)

    def init(self, gas_names):
        self.gas_name = [g.strip().lower() for g in gas_names]
        self.concs = {}
        return ""

    def set_vmr(self, gas, w):
        gas = gas.strip().lower()
        if gas not in self.gas_name:
            self.gas_name.append(gas)
        if np.ndim(w) == 0:
            w = float(w)
            if w < 0.0 or w > 1.0:
                return "ty_gas_concs%set_vmr(): concentrations should be >= 0, <= 1"
        self.concs[gas] = w
        return ""

    def get_vmr(self, gas):
        return self.concs[gas.strip().lower()]

    def _to_c(self, ctx, host=False):
        """-> (array of rrnn_gas_t, keep-alive list)."""
        items = list(self.concs.items())
        arr = (rrnn_gas_t * max(1, len(items)))()
        keep = []
        for i, (name, v) in enumerate(items):
            arr[i].name = name.encode()
            if np.ndim(v) == 0:
                arr[i].ndims = 0; arr[i].value = float(v); arr[i].conc = None
            else:
                if host:
                    t = np.ascontiguousarray(v, np.float32)
                    arr[i].conc = t.ctypes.data
                else:
                    t = _dev(v, ctx)
                    arr[i].conc = t.data_ptr()
                keep.append(t)
                arr[i].ndims = t.ndim
        return arr, len(items), keep


class _kdist_handle:
    def __init__(self, ctx, kd):
        self.h = vp()
        bl = np.ascontiguousarray(kd["band_lims_gpt"], np.int32)
        tot = kd.get("totplnk")
        sol = kd.get("solar_source")
        fp = lambda a: None if a is None else np.ascontiguousarray(a, np.float32).ctypes.data_as(_lib.c_float_p)
        tot_a = None if tot is None else np.ascontiguousarray(tot, np.float32)
        sol_a = None if sol is None else np.ascontiguousarray(sol, np.float32)
        ntemp = 0 if tot_a is None else tot_a.shape[1]
        _lib.check(_lib.lib().rrnn_kdist_create(ctx.h, int(kd["nbnd"]), int(kd["ngpt"]), bl.ctypes.data_as(_lib.c_int_p), ntemp,
                                                fp(tot_a), float(kd.get("temp_ref_min", 0.0)), float(kd.get("totplnk_delta", 1.0)),
                                                fp(sol_a), C.byref(self.h)))
        tabs = [kd.get(k) for k in ("solar_source_quiet", "solar_source_facular", "solar_source_sunspot")]
        if all(t is not None for t in tabs):   # load_ext, rrtmgp/mo_gas_optics_rrtmgp.F90:1317-1325
            arrs = [np.ascontiguousarray(t, np.float32) for t in tabs]
            _lib.check(_lib.lib().rrnn_kdist_set_solar_tables(self.h, *[fp(a) for a in arrs]))
        if kd.get("optimal_angle_fit") is not None:   # (nbnd, 2) == Fortran optimal_angle_fit(2, nbnd), :1163, 1210
            fit = np.ascontiguousarray(kd["optimal_angle_fit"], np.float32)
            if fit.shape != (int(kd["nbnd"]), 2):
                raise RRNNError("gas_optics%load: optimal_angle_fit must be (nbnd, 2)")
            _lib.check(_lib.lib().rrnn_kdist_set_optimal_angle_fit(self.h, fp(fit)))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rrnn_kdist_destroy(self.h)
        except Exception:
            pass


class ty_optical_props:
    """Spectral discretisation base (rte/mo_optical_props.F90:62-96)."""

    def __init__(self):
        self.band2gpt = None
        self.ngpt = 0
        self.nband = 0
        self.name = ""

    def init(self, spectral, name=""):
        if isinstance(spectral, ty_optical_props):
            self.band2gpt, self.ngpt, self.nband = spectral.band2gpt, spectral.ngpt, spectral.nband
            self._kd = getattr(spectral, "_kd", None)
        else:
            self.band2gpt = np.asarray(spectral["band_lims_gpt"], np.int32)
            self.ngpt = int(spectral["ngpt"]); self.nband = int(spectral["nbnd"])
        self.name = name
        return ""

    def get_ngpt(self): return self.ngpt
    def get_nband(self): return self.nband
    def get_band_lims_gpoint(self): return self.band2gpt
    def get_name(self): return self.name


# K5: `clouds.increment(atmos)` with by-band clouds is DEFERRED -- the small by-band arrays are copied and
# remembered on `atmos`; rte_lw / rte_sw then hand them to solvers that add them to the gas optical properties in registers
# (rrnn_rte_lw_clouds / rrnn_rte_sw_clouds), and anything else that looks at atmos.tau / .ssa / .g first applies the increment the
# ordinary way.  Same results either way (tests); FUSE_CLOUD_INCREMENT = False restores the eager pass over (ngpt,nlay,ncol).
FUSE_CLOUD_INCREMENT = True


class _deferred_increment:
    """Mixin for the array carriers: `tau` (and `ssa`) are properties that apply a pending cloud increment before they are seen."""
    _tau = None
    _pending = None

    @property
    def tau(self):
        self._apply_pending()
        return self._tau

    @tau.setter
    def tau(self, v):
        self._pending = None
        self._tau = v

    def _apply_pending(self):
        pend, self._pending = self._pending, None
        if pend is not None:
            msg = pend._increment_now(self)
            if msg != "":
                raise RRNNError(msg)


class ty_optical_props_1scl(_deferred_increment, ty_optical_props):
    def __init__(self):
        super().__init__()
        self.tau = None

    def alloc_1scl(self, ncol, nlay, spectral=None, name="", by_band=False, ctx=None):
        torch = _torch()
        if spectral is not None:
            self.init(spectral, name)
        self.ctx = ctx or getattr(spectral, "ctx", None) or default_context()
        n = self.nband if by_band else self.ngpt
        self.by_band = by_band
        self.tau = torch.empty((ncol, nlay, n), dtype=torch.float32, device=torch.device("cuda", self.ctx.device))
        return ""

    def get_ncol(self): return self._tau.shape[0]
    def get_nlay(self): return self._tau.shape[1]

    def increment(self, op_io):
        """op_io := op_io + self, self given by band (inc_1scalar_by_1scalar_bybnd); deferred into rte_lw where possible."""
        if not getattr(self, "by_band", False):
            op_io.tau += self.tau
            return ""
        if (FUSE_CLOUD_INCREMENT and isinstance(op_io, ty_optical_props_1scl) and not getattr(op_io, "by_band", False)
                and tuple(op_io._tau.shape[:2]) == tuple(self.tau.shape[:2])):
            op_io._apply_pending()
            snap = ty_optical_props_1scl()
            snap.by_band, snap.ctx = True, self.ctx
            with _torch().cuda.stream(op_io.ctx.torch_stream()):
                snap._tau = self.tau.clone()
            op_io._pending = snap
            return ""
        return self._increment_now(op_io)

    def _increment_now(self, op_io):
        try:
            _lib.check(_lib.lib().rrnn_increment_1scl_bybnd(op_io.ctx.h, op_io._kd.h, op_io.get_nlay(), op_io.get_ncol(),
                                                            _ptr(op_io.tau), _ptr(self.tau)))
        except RRNNError as e:
            return str(e)
        return ""


class ty_optical_props_2str(_deferred_increment, ty_optical_props):
    _ssa = None

    def __init__(self):
        super().__init__()
        self.tau = None
        self.ssa = None
        self._g = None
        self.g_is_zero = False

    @property
    def ssa(self):
        self._apply_pending()
        return self._ssa

    @ssa.setter
    def ssa(self, v):
        self._ssa = v

    def alloc_2str(self, ncol, nlay, spectral=None, name="", by_band=False, ctx=None):
        torch = _torch()
        if spectral is not None:
            self.init(spectral, name)
        self.ctx = ctx or getattr(spectral, "ctx", None) or default_context()
        n = self.nband if by_band else self.ngpt
        self.by_band = by_band
        dev = torch.device("cuda", self.ctx.device)
        self.tau = torch.empty((ncol, nlay, n), dtype=torch.float32, device=dev)
        self.ssa = torch.empty((ncol, nlay, n), dtype=torch.float32, device=dev)
        self._g = None
        self.g_is_zero = False
        return ""

    @property
    def g(self):
        """Asymmetry parameter; materialised on first use (gas optics leaves it as an implicit zero)."""
        torch = _torch()
        self._apply_pending()
        if self._g is None:
            with torch.cuda.stream(self.ctx.torch_stream()):
                self._g = (torch.zeros_like(self.tau) if self.g_is_zero else torch.empty_like(self.tau))
        return self._g

    @g.setter
    def g(self, v):
        self._g = v
        self.g_is_zero = False

    def get_ncol(self): return self._tau.shape[0]
    def get_nlay(self): return self._tau.shape[1]

    def delta_scale(self):
        try:
            _lib.check(_lib.lib().rrnn_delta_scale_2str(self.ctx.h, self.tau.numel(), _ptr(self.tau), _ptr(self.ssa), _ptr(self.g)))
        except RRNNError as e:
            return str(e)
        self.g_is_zero = False
        return ""

    def increment(self, op_io):
        """op_io := op_io + self, self given by band (inc_2stream_by_2stream_bybnd); deferred into rte_sw where possible (gas
        optical properties whose g is still the implicit zero of the NN gas optics)."""
        if (FUSE_CLOUD_INCREMENT and getattr(self, "by_band", False) and isinstance(op_io, ty_optical_props_2str)
                and not getattr(op_io, "by_band", False) and op_io._pending is None and op_io.g_is_zero and op_io._g is None
                and tuple(op_io._tau.shape[:2]) == tuple(self.tau.shape[:2])):
            snap = ty_optical_props_2str()
            snap.by_band, snap.ctx = True, self.ctx
            with _torch().cuda.stream(op_io.ctx.torch_stream()):
                snap._tau, snap._ssa, snap._g = self.tau.clone(), self.ssa.clone(), self.g.clone()
            op_io._pending = snap
            return ""
        return self._increment_now(op_io)

    def _increment_now(self, op_io):
        try:
            g1 = op_io.g  # materialises zeros if needed
            _lib.check(_lib.lib().rrnn_increment_2str_bybnd(op_io.ctx.h, op_io._kd.h, op_io.get_nlay(), op_io.get_ncol(),
                                                            _ptr(op_io.tau), _ptr(op_io.ssa), _ptr(g1), _ptr(self.tau),
                                                            _ptr(self.ssa), _ptr(self.g)))
            op_io.g_is_zero = False
        except RRNNError as e:
            return str(e)
        return ""


class ty_source_func_lw(ty_optical_props):
    def __init__(self):
        super().__init__()
        self.lay_source = self.lev_source = self.sfc_source = self.sfc_source_Jac = None

    def alloc(self, ncol, nlay, spectral=None, ctx=None):
        torch = _torch()
        if spectral is not None:
            self.init(spectral)
        self.ctx = ctx or getattr(spectral, "ctx", None) or default_context()
        dev = torch.device("cuda", self.ctx.device)
        G = self.ngpt
        self.lay_source = torch.empty((ncol, nlay, G), dtype=torch.float32, device=dev)
        self.lev_source = torch.empty((ncol, nlay + 1, G), dtype=torch.float32, device=dev)
        self.sfc_source = torch.empty((ncol, G), dtype=torch.float32, device=dev)
        self.sfc_source_Jac = torch.empty((ncol, G), dtype=torch.float32, device=dev)
        return ""

    def get_ncol(self): return self.lay_source.shape[0]
    def get_nlay(self): return self.lay_source.shape[1]


class ty_fluxes_broadband:
    """Output holder: the caller associates the arrays it wants (all (ncol, nlay+1))."""

    def __init__(self, flux_up=None, flux_dn=None, flux_net=None, flux_dn_dir=None):
        self.flux_up, self.flux_dn, self.flux_net, self.flux_dn_dir = flux_up, flux_dn, flux_net, flux_dn_dir

    def are_desired(self):
        return any(v is not None for v in (self.flux_up, self.flux_dn, self.flux_net, self.flux_dn_dir))


class ty_fluxes_flexible(ty_fluxes_broadband):
    """ty_fluxes_flexible (rte/mo_fluxes.F90:52-67): broadband fluxes plus, when associated, the g-point fluxes
    gpt_flux_up / gpt_flux_dn (ncol, nlay+1, ngpt)."""

    def __init__(self, flux_up=None, flux_dn=None, flux_net=None, flux_dn_dir=None, gpt_flux_up=None, gpt_flux_dn=None,
                 gpt_flux_dn_dir=None):
        super().__init__(flux_up, flux_dn, flux_net, flux_dn_dir)
        self.gpt_flux_up, self.gpt_flux_dn, self.gpt_flux_dn_dir = gpt_flux_up, gpt_flux_dn, gpt_flux_dn_dir


class ty_fluxes_byband(ty_fluxes_broadband):
    """ty_fluxes_byband (extensions/mo_fluxes_byband.F90:31-40): broadband fluxes plus, when associated, the by-band fluxes
    bnd_flux_up / bnd_flux_dn / bnd_flux_net / bnd_flux_dn_dir, (ncol, nlay+1, nband) in this fork's band-fastest layout."""

    def __init__(self, flux_up=None, flux_dn=None, flux_net=None, flux_dn_dir=None, bnd_flux_up=None, bnd_flux_dn=None,
                 bnd_flux_net=None, bnd_flux_dn_dir=None):
        super().__init__(flux_up, flux_dn, flux_net, flux_dn_dir)
        self.bnd_flux_up, self.bnd_flux_dn, self.bnd_flux_net, self.bnd_flux_dn_dir = bnd_flux_up, bnd_flux_dn, bnd_flux_net, bnd_flux_dn_dir

    def _bnd_desired(self):
        return any(v is not None for v in (self.bnd_flux_up, self.bnd_flux_dn, self.bnd_flux_net, self.bnd_flux_dn_dir))

    def are_desired(self):
        return self._bnd_desired() or super().are_desired()

    def reduce(self, gpt_flux_up, gpt_flux_dn, spectral_disc, top_at_1, gpt_flux_dn_dir=None):
        """reduce_byband (:41-131), by-band part: g-point fluxes (ncol, nlay+1, ngpt) on the device -> the associated by-band
        arrays.  (The broadband part of reduce is done inside the solvers' fused broadband sums.)"""
        ncol, nlev, ngpt = tuple(gpt_flux_up.shape)
        nbnd = spectral_disc.get_nband()
        if tuple(gpt_flux_dn.shape) != (ncol, nlev, ngpt):
            return "reduce: gpt_flux_dn array incorrectly sized"
        if gpt_flux_dn_dir is not None and tuple(gpt_flux_dn_dir.shape) != (ncol, nlev, ngpt):
            return "reduce: gpt_flux_dn_dir array incorrectly sized"
        if ngpt != spectral_disc.get_ngpt():
            return "reduce: spectral discretization and g-point flux arrays have differing number of g-points"
        err = ""
        for nm in ("bnd_flux_up", "bnd_flux_dn"):
            v = getattr(self, nm)
            if v is not None and tuple(v.shape) != (ncol, nlev, nbnd):
                err = f"reduce: {nm} array incorrectly sized (can't compute net flux either)"
        if self.bnd_flux_dn_dir is not None and tuple(self.bnd_flux_dn_dir.shape) != (ncol, nlev, nbnd):
            err = "reduce: bnd_flux_dn_dir array incorrectly sized"
        if self.bnd_flux_net is not None and tuple(self.bnd_flux_net.shape) != (ncol, nlev, nbnd):
            err = "reduce: bnd_flux_net array incorrectly sized (can't compute net flux either)"
        if err:
            return err
        if self.bnd_flux_dn_dir is not None and gpt_flux_dn_dir is None:
            return "reduce: requesting bnd_flux_dn_dir but direct flux hasn't been supplied"
        kd = getattr(spectral_disc, "_kd", None)
        if kd is None:
            return "reduce: spectral discretization carries no band limits"
        ctx = getattr(spectral_disc, "ctx", None) or default_context()
        lib = _lib.lib()
        try:
            for out, src in ((self.bnd_flux_up, gpt_flux_up), (self.bnd_flux_dn, gpt_flux_dn), (self.bnd_flux_dn_dir, gpt_flux_dn_dir)):
                if out is not None:
                    _lib.check(lib.rrnn_sum_byband(ctx.h, kd.h, nlev, ncol, _ptr(src), _ptr(out)))
            if self.bnd_flux_net is not None:
                if self.bnd_flux_dn is not None and self.bnd_flux_up is not None:   # net_byband_precalc
                    _lib.check(lib.rrnn_net_flux(ctx.h, ncol * nlev * nbnd, _ptr(self.bnd_flux_dn), _ptr(self.bnd_flux_up),
                                                 _ptr(self.bnd_flux_net)))
                else:                                                               # net_byband_full
                    _lib.check(lib.rrnn_net_byband(ctx.h, kd.h, nlev, ncol, _ptr(gpt_flux_dn), _ptr(gpt_flux_up),
                                                   _ptr(self.bnd_flux_net)))
        except RRNNError as e:
            return str(e)
        return ""


# ty_fluxes_byband from the tuned solvers (True, default) or the reference's own route -- g-point fluxes from the general kernels, then
# sum_byband in g-point order, bit-exact against the oracle's serial sums (False)
BYBAND_FROM_SOLVER = True


def _byband_from_solver(fluxes, optical_props, sw, call):
    """ty_fluxes_byband straight from the tuned solvers (rrnn_rte_{lw,sw}_byband: the per-level sums stop at a band on their way to
    the broadband sum) -- no g-point fluxes, no general kernel.  `call(up, dn, dir, bnd_up, bnd_dn, bnd_dir)` runs the C entry point.
    Returns None when the packed solver does not take the case (the caller then goes the general way), else the error message."""
    torch = _torch()
    ctx = optical_props.ctx
    ncol, nlev, nbnd = optical_props.get_ncol(), optical_props.get_nlay() + 1, optical_props.nband
    dev = optical_props.tau.device
    if not BYBAND_FROM_SOLVER:
        return None
    for nm in ("flux_up", "flux_dn", "flux_net", "flux_dn_dir"):
        v = getattr(fluxes, nm)
        if v is not None and tuple(v.shape) != (ncol, nlev):
            return f"reduce: {nm} array incorrectly sized"
    for nm, msg in (("bnd_flux_up", "reduce: bnd_flux_up array incorrectly sized (can't compute net flux either)"),
                    ("bnd_flux_dn", "reduce: bnd_flux_dn array incorrectly sized (can't compute net flux either)"),
                    ("bnd_flux_dn_dir", "reduce: bnd_flux_dn_dir array incorrectly sized"),
                    ("bnd_flux_net", "reduce: bnd_flux_net array incorrectly sized (can't compute net flux either)")):
        v = getattr(fluxes, nm)
        if v is not None and tuple(v.shape) != (ncol, nlev, nbnd):
            return msg
    if not sw and fluxes.bnd_flux_dn_dir is not None:
        return "reduce: requesting bnd_flux_dn_dir but direct flux hasn't been supplied"

    def pick(v, *shape):
        if v is not None:
            return v
        with torch.cuda.stream(ctx.torch_stream()):
            return torch.empty(shape, dtype=torch.float32, device=dev)
    up, dn = pick(fluxes.flux_up, ncol, nlev), pick(fluxes.flux_dn, ncol, nlev)
    dr = pick(fluxes.flux_dn_dir, ncol, nlev) if sw else None
    bup, bdn = pick(fluxes.bnd_flux_up, ncol, nlev, nbnd), pick(fluxes.bnd_flux_dn, ncol, nlev, nbnd)
    bdr = pick(fluxes.bnd_flux_dn_dir, ncol, nlev, nbnd) if sw else None
    try:
        call(up, dn, dr, bup, bdn, bdr)
        if fluxes.flux_net is not None:
            _lib.check(_lib.lib().rrnn_net_flux(ctx.h, ncol * nlev, _ptr(dn), _ptr(up), _ptr(fluxes.flux_net)))
        if fluxes.bnd_flux_net is not None:
            _lib.check(_lib.lib().rrnn_net_flux(ctx.h, ncol * nlev * nbnd, _ptr(bdn), _ptr(bup), _ptr(fluxes.bnd_flux_net)))
    except RRNNError as e:
        if "by-band fluxes from the packed solver need" in str(e):
            return None
        return str(e)
    return ""


def _solve_and_reduce(core, fluxes, optical_props, sw):
    """Wraps a solver call for the flux types whose outputs the fused broadband kernels do not produce themselves: flux_net
    (ty_fluxes_broadband%reduce, rte/mo_fluxes.F90: net_broadband_precalc) and ty_fluxes_byband's by-band arrays, which need
    the g-point fluxes (general kernels) reduced by band afterwards."""
    byband = isinstance(fluxes, ty_fluxes_byband) and fluxes._bnd_desired()
    if not byband and fluxes.flux_net is None:
        return core(fluxes)
    torch = _torch()
    ctx = optical_props.ctx
    ncol, nlev, ngpt = optical_props.get_ncol(), optical_props.get_nlay() + 1, optical_props.ngpt
    dev = optical_props.tau.device

    def new(*shape):   # temporaries belong to the context's stream (the caching allocator re-uses them in stream order)
        with torch.cuda.stream(ctx.torch_stream()):
            return torch.empty(shape, dtype=torch.float32, device=dev)

    for nm in ("flux_up", "flux_dn", "flux_net", "flux_dn_dir"):
        v = getattr(fluxes, nm)
        if v is not None and tuple(v.shape) != (ncol, nlev):
            return f"reduce: {nm} array incorrectly sized"
    pick = lambda v: v if v is not None else new(ncol, nlev)
    inner = ty_fluxes_flexible(pick(fluxes.flux_up), pick(fluxes.flux_dn), None,
                               pick(fluxes.flux_dn_dir) if sw else None,
                               getattr(fluxes, "gpt_flux_up", None), getattr(fluxes, "gpt_flux_dn", None),
                               getattr(fluxes, "gpt_flux_dn_dir", None))
    if byband:
        if inner.gpt_flux_up is None: inner.gpt_flux_up = new(ncol, nlev, ngpt)
        if inner.gpt_flux_dn is None: inner.gpt_flux_dn = new(ncol, nlev, ngpt)
        if sw and inner.gpt_flux_dn_dir is None: inner.gpt_flux_dn_dir = new(ncol, nlev, ngpt)
        if not sw and fluxes.bnd_flux_dn_dir is not None:
            return "reduce: requesting bnd_flux_dn_dir but direct flux hasn't been supplied"
    err = core(inner)
    if err:
        return err
    if byband:
        err = fluxes.reduce(inner.gpt_flux_up, inner.gpt_flux_dn, optical_props, None, inner.gpt_flux_dn_dir if sw else None)
        if err:
            return err
    if fluxes.flux_net is not None:
        try:
            _lib.check(_lib.lib().rrnn_net_flux(ctx.h, ncol * nlev, _ptr(inner.flux_dn), _ptr(inner.flux_up), _ptr(fluxes.flux_net)))
        except RRNNError as e:
            return str(e)
    return ""


class ty_gas_optics_rrtmgp(ty_optical_props):
    """The NN path of the RRTMGP gas optics.  `load` takes the spectral tables of the k-distribution
    (rte_rrtmgp_nn_b200.spectral.make_kdist / synthetic_kdist_*)."""

    def __init__(self, ctx=None):
        super().__init__()
        self.ctx = ctx or default_context()
        self.kd = None
        self._kd = None

    def load(self, kd):
        self.kd = dict(kd)
        self.init(kd, "ty_gas_optics_rrtmgp")
        try:
            self._kd = _kdist_handle(self.ctx, self.kd)
        except RRNNError as e:
            return str(e)
        return ""

    def source_is_internal(self):
        return self.kd.get("totplnk") is not None

    def source_is_external(self):
        return self.kd.get("solar_source") is not None

    def get_press_min(self): return self.kd["press_ref_min"]
    def get_press_max(self): return self.kd["press_ref_max"]
    def get_temp_min(self): return self.kd["temp_ref_min"]
    def get_temp_max(self): return self.kd["temp_ref_max"]

    def set_tsi(self, tsi):
        try:
            _lib.check(_lib.lib().rrnn_kdist_set_tsi(self._kd.h, float(tsi)))
        except RRNNError as e:
            return str(e)
        return ""

    def set_solar_variability(self, mg_index, sb_index, tsi=None):
        """set_solar_variability (rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1095); needs solar_source_quiet / _facular / _sunspot
        among the tables given to load()."""
        try:
            _lib.check(_lib.lib().rrnn_kdist_set_solar_variability(self._kd.h, float(mg_index), float(sb_index),
                                                                   int(tsi is not None), float(0.0 if tsi is None else tsi)))
        except RRNNError as e:
            return str(e)
        return ""

    def get_solar_source(self):
        out = np.empty(self.ngpt, np.float32)
        _lib.check(_lib.lib().rrnn_kdist_get_solar_source(self._kd.h, out.ctypes.data_as(_lib.c_float_p)))
        return out

    def gpoints_are_equal(self, other):
        return self.ngpt == other.get_ngpt() and np.array_equal(self.band2gpt, other.get_band_lims_gpoint())

    def compute_optimal_angles(self, optical_props, optimal_angles):
        """compute_optimal_angles (rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758): optimal_angles is a device tensor (ncol, ngpt),
        the layout rte_lw's lw_Ds takes."""
        if not self.gpoints_are_equal(optical_props):
            return "gas_optics%compute_optimal_angles: optical_props has different spectral discretization than gas_optics"
        ncol, nlay = optical_props.get_ncol(), optical_props.get_nlay()
        if tuple(optimal_angles.shape) != (ncol, self.ngpt):
            return "gas_optics%compute_optimal_angles: optimal_angles different dimension (ncol)"
        try:
            _lib.check(_lib.lib().rrnn_compute_optimal_angles(self.ctx.h, self._kd.h, nlay, ncol, _ptr(optical_props.tau),
                                                              _ptr(optimal_angles)))
        except RRNNError as e:
            return str(e)
        return ""

    def gas_optics(self, play, plev, tlay, *args, col_dry=None, tlev=None, neural_nets=None):
        """LW: gas_optics(play, plev, tlay, tsfc, gas_desc, optical_props, sources [,col_dry][,tlev][,neural_nets])
           SW: gas_optics(play, plev, tlay, gas_desc, optical_props, toa_src [,col_dry][,neural_nets])."""
        if neural_nets is None:
            return "gas_optics(): only the neural-network path (neural_nets=) is implemented; the LUT path is out of scope"
        if col_dry is not None:
            return "gas_optics(): optional col_dry is not supported on the neural-network path (as in the reference)"
        ctx = self.ctx
        play, plev, tlay = _dev(play, ctx), _dev(plev, ctx), _dev(tlay, ctx)
        ncol, nlay = play.shape
        lib = _lib.lib()
        models = (vp * 2)(*[n.h for n in neural_nets][:2])
        lw = self.source_is_internal()
        # mo_gas_optics_rrtmgp.F90:287-315 (LW), 474-494 (SW): the same checks, the same messages, off unless configured
        if check_extents:
            err = ""
            if tuple(play.shape) != (ncol, nlay): err = "gas_optics(): array play has wrong size"
            if tuple(tlay.shape) != (ncol, nlay): err = "gas_optics(): array tlay has wrong size"
            if tuple(plev.shape) != (ncol, nlay + 1): err = "gas_optics(): array plev has wrong size"
            if lw and np.shape(args[0]) != (ncol,): err = "gas_optics(): array tsfc has wrong size"
            if lw and tlev is not None and tuple(np.shape(tlev)) != (ncol, nlay + 1): err = "gas_optics(): array tlev has wrong size"
            if err:
                return err
        if check_values:
            err = ""
            pmin, pmax = self.kd.get("press_ref_min", 1.00518357), self.kd.get("press_ref_max", 109663.31)
            tmin, tmax = self.kd.get("temp_ref_min", 160.0), self.kd.get("temp_ref_max", 355.0)
            if _outside(play, pmin, pmax): err = "gas_optics(): array play has values outside range"
            if bool((plev < 0).any()): err = "gas_optics(): array plev has values outside range"
            if _outside(tlay, tmin, tmax): err = "gas_optics(): array tlay has values outside range"
            if lw and _outside(_dev(args[0], ctx), tmin, tmax): err = "gas_optics(): array tsfc has values outside range"
            if lw and tlev is not None and _outside(_dev(tlev, ctx), tmin, tmax): err = "gas_optics(): array tlev has values outside range"
            if err:
                return err
        try:
            if lw:
                tsfc, gas_desc, optical_props, sources = args
                tsfc = _dev(tsfc, ctx)
                tlev_d = _dev(tlev, ctx)
                gases, ngas, keep = gas_desc._to_c(ctx)
                if sources.get_ncol() != ncol or sources.get_nlay() != nlay or sources.ngpt != self.ngpt:
                    return "gas_optics%gas_optics: source function arrays inconsistently sized"
                optical_props._kd = self._kd
                optical_props._pending = None      # everything in optical_props is overwritten
                _lib.check(lib.rrnn_gas_optics_lw(ctx.h, self._kd.h, models, len(neural_nets), ncol, nlay, _ptr(play), _ptr(plev),
                                                  _ptr(tlay), _ptr(tsfc), gases, ngas, _ptr(tlev_d), _ptr(optical_props.tau),
                                                  _ptr(sources.lay_source), _ptr(sources.lev_source), _ptr(sources.sfc_source),
                                                  _ptr(sources.sfc_source_Jac)))
            else:
                gas_desc, optical_props, toa_src = args
                gases, ngas, keep = gas_desc._to_c(ctx)
                optical_props._kd = self._kd
                optical_props._pending = None      # everything in optical_props is overwritten
                if isinstance(optical_props, ty_optical_props_2str):
                    optical_props._g = None
                    optical_props.g_is_zero = True   # g(:,:,:) = 0, mo_gas_optics_rrtmgp.F90:560-567 (kept implicit)
                    ssa_p = _ptr(optical_props.ssa)
                else:
                    ssa_p = None
                if tuple(toa_src.shape) != (ncol, self.ngpt):
                    return "gas_optics(): array toa_src has wrong size"
                _lib.check(lib.rrnn_gas_optics_sw(ctx.h, self._kd.h, models, ncol, nlay, _ptr(play), _ptr(plev), _ptr(tlay), gases,
                                                  ngas, _ptr(optical_props.tau), ssa_p, None, _ptr(toa_src)))
        except RRNNError as e:
            return str(e)
        return ""

    def gas_optics_compact(self, play, plev, tlay, tsfc, gas_desc, tau, pfrac, planck_lay, planck_lev, sfc_source, sfc_source_Jac,
                           tlev=None, neural_nets=None):
        """LW gas optics with the sources left factored (rrnn_gas_optics_lw_compact, no reference counterpart): device
        tensors tau, pfrac (ncol,nlay,ngpt), planck_lay (ncol,nlay,16), planck_lev (ncol,nlay+1,16), sfc_source[_Jac] (ncol,ngpt)."""
        ctx = self.ctx
        play, plev, tlay, tsfc = _dev(play, ctx), _dev(plev, ctx), _dev(tlay, ctx), _dev(tsfc, ctx)
        ncol, nlay = play.shape
        models = (vp * 2)(*[n.h for n in neural_nets][:2])
        try:
            tlev_d = _dev(tlev, ctx)
            gases, ngas, keep = gas_desc._to_c(ctx)
            _lib.check(_lib.lib().rrnn_gas_optics_lw_compact(ctx.h, self._kd.h, models, len(neural_nets), ncol, nlay, _ptr(play), _ptr(plev),
                                                             _ptr(tlay), _ptr(tsfc), gases, ngas, _ptr(tlev_d), _ptr(tau), _ptr(pfrac),
                                                             _ptr(planck_lay), _ptr(planck_lev), _ptr(sfc_source), _ptr(sfc_source_Jac)))
        except RRNNError as e:
            return str(e)
        return ""


def rte_lw(optical_props, top_at_1, sources, sfc_emis, fluxes, inc_flux=None, n_gauss_angles=None, use_2stream=None,
           lw_Ds=None, flux_up_Jac=None, flux_dn_Jac=None):
    """rte_lw (rte/mo_rte_lw.F90:60-64); sfc_emis is (ncol, nband).  ty_optical_props_1scl: no-scattering solution (the tuned
    kernels; the general kernel when lw_Ds, flux_up_Jac or g-point fluxes are asked for).  ty_optical_props_2str: the
    re-scaled solution (:363-384), or lw_solver_2stream with use_2stream=True (:346-361)."""
    if not fluxes.are_desired():
        return "rte_lw: no space allocated for fluxes"
    if fluxes.flux_net is not None or (isinstance(fluxes, ty_fluxes_byband) and fluxes._bnd_desired()):
        if (isinstance(fluxes, ty_fluxes_byband) and fluxes._bnd_desired() and isinstance(optical_props, ty_optical_props_1scl)
                and not use_2stream and lw_Ds is None and flux_up_Jac is None and flux_dn_Jac is None
                and (n_gauss_angles is None or 1 <= int(n_gauss_angles) <= 4)):
            ctx = optical_props.ctx
            ncol, nlay = optical_props.get_ncol(), optical_props.get_nlay()
            emis_d = _dev(sfc_emis, ctx)
            if tuple(emis_d.shape) != (ncol, optical_props.nband):
                return "rte_lw: sfc_emis inconsistently sized"
            inc_d = _dev(inc_flux, ctx)
            nang = 1 if n_gauss_angles is None else int(n_gauss_angles)
            err = _byband_from_solver(fluxes, optical_props, False, lambda up, dn, dr, bup, bdn, bdr: _lib.check(_lib.lib().rrnn_rte_lw_byband(
                ctx.h, optical_props._kd.h, nlay, ncol, int(bool(top_at_1)), nang, _ptr(inc_d), _ptr(optical_props.tau), _ptr(sources.lay_source),
                _ptr(sources.lev_source), _ptr(sources.sfc_source), _ptr(emis_d), _ptr(up), _ptr(dn), _ptr(bup), _ptr(bdn))))
            if err is not None:
                return err
        return _solve_and_reduce(lambda f: rte_lw(optical_props, top_at_1, sources, sfc_emis, f, inc_flux, n_gauss_angles, use_2stream,
                                                  lw_Ds, flux_up_Jac, flux_dn_Jac), fluxes, optical_props, sw=False)
    two = isinstance(optical_props, ty_optical_props_2str)
    if not two and not isinstance(optical_props, ty_optical_props_1scl):
        return "rte_lw: lw_solver(...ty_optical_props_nstr...) not yet implemented"
    if use_2stream and not two:
        return "rte_lw: can't use two-stream methods with only absorption optical depth"
    if use_2stream and (flux_up_Jac is not None or flux_dn_Jac is not None):
        return "rte_lw: can't provide Jacobian of fluxes w.r.t surface temperature with 2-stream"
    if use_2stream:
        if n_gauss_angles is not None:
            return "rte_lw: n_gauss_angles not compatible with use_2stream"
        if lw_Ds is not None:
            return "rte_lw: lw_Ds not valid input for _2str class"
        ctx = optical_props.ctx
        ncol, nlay = optical_props.get_ncol(), optical_props.get_nlay()
        sfc_emis = _dev(sfc_emis, ctx)
        if tuple(sfc_emis.shape) != (ncol, optical_props.nband):
            return "rte_lw: sfc_emis inconsistently sized"
        try:
            _lib.check(_lib.lib().rrnn_rte_lw_2stream(
                ctx.h, optical_props._kd.h, nlay, ncol, int(bool(top_at_1)), _ptr(_dev(inc_flux, ctx)), _ptr(optical_props.tau),
                _ptr(optical_props.ssa), _ptr(optical_props.g), _ptr(sources.lev_source), _ptr(sources.sfc_source), _ptr(sfc_emis),
                _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn), _ptr(getattr(fluxes, "gpt_flux_up", None)),
                _ptr(getattr(fluxes, "gpt_flux_dn", None))))
        except RRNNError as e:
            return str(e)
        return ""
    if flux_dn_Jac is not None:
        return "rte_lw: flux_dn_Jac is not computed (as in the reference)"
    if lw_Ds is not None and two:
        return "rte_lw: lw_Ds not valid input for _2str class"
    if lw_Ds is not None and n_gauss_angles is not None:
        return "rte_lw: providing lw_Ds incompatible with specifying n_gauss_angles"
    ctx = optical_props.ctx
    gpt_up, gpt_dn = getattr(fluxes, "gpt_flux_up", None), getattr(fluxes, "gpt_flux_dn", None)
    if two or lw_Ds is not None or flux_up_Jac is not None or gpt_up is not None:
        nang = 1 if n_gauss_angles is None else int(n_gauss_angles)
        ncol, nlay = optical_props.get_ncol(), optical_props.get_nlay()
        sfc_emis = _dev(sfc_emis, ctx)
        if tuple(sfc_emis.shape) != (ncol, optical_props.nband):
            return "rte_lw: sfc_emis inconsistently sized"
        if lw_Ds is not None:
            lw_Ds = _dev(lw_Ds, ctx)
            if tuple(lw_Ds.shape) != (ncol, optical_props.ngpt):
                return "rte_lw: lw_Ds inconsistently sized"
            if bool((lw_Ds < 1.0).any()):
                return "rte_lw: one or more values of lw_Ds < 1."
        if flux_up_Jac is not None and tuple(flux_up_Jac.shape) != (ncol, nlay + 1):
            return "rte_lw: flux Jacobian inconsistently sized"
        if two:
            g = optical_props.g  # materialises zeros when g is implicit
        try:
            _lib.check(_lib.lib().rrnn_rte_lw_ext(
                ctx.h, optical_props._kd.h, nlay, ncol, int(bool(top_at_1)), nang, _ptr(_dev(inc_flux, ctx)), _ptr(optical_props.tau),
                _ptr(optical_props.ssa) if two else None, _ptr(g) if two else None, _ptr(sources.lay_source),
                _ptr(sources.lev_source), _ptr(sources.sfc_source), _ptr(sfc_emis), _ptr(lw_Ds),
                _ptr(sources.sfc_source_Jac) if flux_up_Jac is not None else None, _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn),
                _ptr(flux_up_Jac), _ptr(gpt_up), _ptr(gpt_dn)))
        except RRNNError as e:
            return str(e)
        return ""
    nang = 1 if n_gauss_angles is None else int(n_gauss_angles)
    if nang > 4:
        return "rte_lw: asking for too many quadrature points for no-scattering calculation"
    if nang < 1:
        return "rte_lw: have to ask for at least one quadrature point for no-scattering calculation"
    ncol, nlay = optical_props.get_ncol(), optical_props.get_nlay()
    sfc_emis = _dev(sfc_emis, ctx)
    if tuple(sfc_emis.shape) != (ncol, optical_props.nband):
        return "rte_lw: sfc_emis inconsistently sized"
    inc = _dev(inc_flux, ctx)
    pend = optical_props._pending
    if pend is not None:     # clouds whose increment is still pending: added to tau inside the solver (K5)
        try:
            _lib.check(_lib.lib().rrnn_rte_lw_clouds(ctx.h, optical_props._kd.h, nlay, ncol, int(bool(top_at_1)), nang, _ptr(inc),
                                                     _ptr(optical_props._tau), _ptr(sources.lay_source), _ptr(sources.lev_source),
                                                     _ptr(sources.sfc_source), _ptr(sfc_emis), _ptr(pend._tau), _ptr(fluxes.flux_up),
                                                     _ptr(fluxes.flux_dn)))
            return ""
        except RRNNError as e:
            if "not taken by the packed solver" not in str(e):
                return str(e)         # (otherwise: apply the increment the ordinary way -- optical_props.tau does -- and go on)
    try:
        _lib.check(_lib.lib().rrnn_rte_lw(ctx.h, optical_props._kd.h, nlay, ncol, int(bool(top_at_1)), nang, _ptr(inc),
                                          _ptr(optical_props.tau), _ptr(sources.lay_source), _ptr(sources.lev_source),
                                          _ptr(sources.sfc_source), _ptr(sfc_emis), _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn)))
    except RRNNError as e:
        return str(e)
    return ""


def rte_sw(atmos, top_at_1, mu0, inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt, fluxes, inc_flux_dif=None):
    """rte_sw (rte/mo_rte_sw.F90:48-52) for ty_optical_props_2str; albedos are per g-point (ncol, ngpt)."""
    if not fluxes.are_desired():
        return "rte_sw: no space allocated for fluxes"
    if not isinstance(atmos, ty_optical_props_2str):
        return "rte_sw: only ty_optical_props_2str (two-stream) is implemented"
    if fluxes.flux_net is not None or (isinstance(fluxes, ty_fluxes_byband) and fluxes._bnd_desired()):
        if isinstance(fluxes, ty_fluxes_byband) and fluxes._bnd_desired():
            ctx = atmos.ctx
            ncol, nlay, ngpt = atmos.get_ncol(), atmos.get_nlay(), atmos.ngpt
            d = [_dev(v, ctx) for v in (mu0, inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt)]
            if tuple(d[0].shape) == (ncol,) and all(tuple(v.shape) == (ncol, ngpt) for v in d[1:]):
                g_p = None if atmos.g_is_zero and atmos._g is None and atmos._pending is None else _ptr(atmos.g)
                dif = _dev(inc_flux_dif, ctx)
                err = _byband_from_solver(fluxes, atmos, True, lambda up, dn, dr, bup, bdn, bdr: _lib.check(_lib.lib().rrnn_rte_sw_byband(
                    ctx.h, atmos._kd.h, nlay, ncol, int(bool(top_at_1)), _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(d[3]), _ptr(dif), _ptr(atmos.tau),
                    _ptr(atmos.ssa), g_p, _ptr(up), _ptr(dn), _ptr(dr), _ptr(bup), _ptr(bdn), _ptr(bdr))))
                if err is not None:
                    return err
        return _solve_and_reduce(lambda f: rte_sw(atmos, top_at_1, mu0, inc_flux, sfc_alb_dir_gpt, sfc_alb_dif_gpt, f, inc_flux_dif),
                                 fluxes, atmos, sw=True)
    ctx = atmos.ctx
    ncol, nlay, ngpt = atmos.get_ncol(), atmos.get_nlay(), atmos.ngpt
    mu0, inc_flux = _dev(mu0, ctx), _dev(inc_flux, ctx)
    a_dir, a_dif = _dev(sfc_alb_dir_gpt, ctx), _dev(sfc_alb_dif_gpt, ctx)
    if tuple(mu0.shape) != (ncol,):
        return "rte_sw: mu0 inconsistently sized"
    if tuple(inc_flux.shape) != (ncol, ngpt):
        return "rte_sw: inc_flux inconsistently sized"
    if tuple(a_dir.shape) != (ncol, ngpt):
        return "rte_sw: sfc_alb_dir inconsistently sized"
    if tuple(a_dif.shape) != (ncol, ngpt):
        return "rte_sw: sfc_alb_dif inconsistently sized"
    gpt = [getattr(fluxes, k, None) for k in ("gpt_flux_up", "gpt_flux_dn", "gpt_flux_dn_dir")]
    pend = atmos._pending
    if pend is not None and all(v is None for v in gpt):   # clouds whose increment is still pending: folded into the solver (K5)
        try:
            _lib.check(_lib.lib().rrnn_rte_sw_clouds(ctx.h, atmos._kd.h, nlay, ncol, int(bool(top_at_1)), _ptr(mu0), _ptr(inc_flux),
                                                     _ptr(a_dir), _ptr(a_dif), _ptr(_dev(inc_flux_dif, ctx)), _ptr(atmos._tau),
                                                     _ptr(atmos._ssa), _ptr(pend._tau), _ptr(pend._ssa), _ptr(pend._g),
                                                     _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn), _ptr(fluxes.flux_dn_dir)))
            return ""
        except RRNNError as e:
            if "not taken by the packed solver" not in str(e):
                return str(e)
    g_p = None if atmos.g_is_zero and atmos._g is None and atmos._pending is None else _ptr(atmos.g)
    if any(v is not None for v in gpt):  # ty_fluxes_flexible with g-point fluxes: the general kernel
        if any(v is None for v in gpt):
            return "rte_sw: gpt_flux_up, gpt_flux_dn and gpt_flux_dn_dir must all be associated"
        if any(tuple(v.shape) != (ncol, nlay + 1, ngpt) for v in gpt):
            return "rte_sw: g-point flux arrays inconsistently sized"
        try:
            _lib.check(_lib.lib().rrnn_sw_solver_2stream_ext(ctx.h, ngpt, nlay, ncol, int(bool(top_at_1)), _ptr(inc_flux),
                                                             _ptr(_dev(inc_flux_dif, ctx)), _ptr(atmos.tau), _ptr(atmos.ssa), g_p, _ptr(mu0),
                                                             _ptr(a_dir), _ptr(a_dif), _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn),
                                                             _ptr(fluxes.flux_dn_dir), *[_ptr(v) for v in gpt]))
        except RRNNError as e:
            return str(e)
        return ""
    try:
        _lib.check(_lib.lib().rrnn_rte_sw(ctx.h, ngpt, nlay, ncol, int(bool(top_at_1)), _ptr(mu0), _ptr(inc_flux), _ptr(a_dir),
                                          _ptr(a_dif), _ptr(_dev(inc_flux_dif, ctx)), _ptr(atmos.tau), _ptr(atmos.ssa), g_p,
                                          _ptr(fluxes.flux_up), _ptr(fluxes.flux_dn), _ptr(fluxes.flux_dn_dir)))
    except RRNNError as e:
        return str(e)
    return ""


class ty_cloud_optics(ty_optical_props):
    def __init__(self, ctx=None):
        super().__init__()
        self.ctx = ctx or default_context()
        self.h = vp()
        self.icergh = 0

    def load(self, band_lims_wvn, radliq_lwr, radliq_upr, radliq_fac, radice_lwr, radice_upr, radice_fac,
             lut_extliq, lut_ssaliq, lut_asyliq, lut_extice, lut_ssaice, lut_asyice, ice_roughness=2):
        """load_lut; tables as stored in the coefficient files: liq [nbnd][nsize], ice [nrough][nbnd][nsize]."""
        self.nband = self.ngpt = int(np.shape(lut_extliq)[0])
        self.band2gpt = np.array([[b + 1, b + 1] for b in range(self.nband)], np.int32)
        self.icergh = int(ice_roughness)
        r = self.icergh - 1
        fa = lambda a: np.ascontiguousarray(a, np.float32)
        t = [fa(lut_extliq), fa(lut_ssaliq), fa(lut_asyliq), fa(np.asarray(lut_extice)[r]), fa(np.asarray(lut_ssaice)[r]),
             fa(np.asarray(lut_asyice)[r])]
        self.tables = dict(extliq=t[0], ssaliq=t[1], asyliq=t[2], extice=t[3], ssaice=t[4], asyice=t[5],
                           liq_nsteps=t[0].shape[1], ice_nsteps=t[3].shape[1], radliq_lwr=float(radliq_lwr),
                           radice_lwr=float(radice_lwr),
                           liq_step_size=float(np.float32(radliq_upr - radliq_lwr) / np.float32(t[0].shape[1] - 1)),
                           ice_step_size=float(np.float32(radice_upr - radice_lwr) / np.float32(t[3].shape[1] - 1)))
        p = [a.ctypes.data_as(_lib.c_float_p) for a in t]
        try:
            _lib.check(_lib.lib().rrnn_cloud_lut_create(self.ctx.h, self.nband, t[0].shape[1], t[3].shape[1], float(radliq_lwr),
                                                        float(radliq_upr), float(radice_lwr), float(radice_upr), *p,
                                                        C.byref(self.h)))
        except RRNNError as e:
            return str(e)
        return ""

    def load_pade(self, band_lims_wvn, pade_extliq, pade_ssaliq, pade_asyliq, pade_extice, pade_ssaice, pade_asyice,
                  pade_sizreg_extliq, pade_sizreg_ssaliq, pade_sizreg_asyliq, pade_sizreg_extice, pade_sizreg_ssaice,
                  pade_sizreg_asyice, ice_roughness=2):
        """load_pade (extensions/cloud_optics/mo_cloud_optics.F90:178-262); arrays as stored in the coefficient files: liq
        (ncoeff, nsizereg, nbnd), ice (nrough, ncoeff, nsizereg, nbnd), size-regime bounds (nbound)."""
        fa = lambda a: np.ascontiguousarray(a, np.float32)
        self.icergh = int(ice_roughness)
        r = self.icergh - 1
        t = [fa(pade_extliq), fa(pade_ssaliq), fa(pade_asyliq), fa(np.asarray(pade_extice)[r]), fa(np.asarray(pade_ssaice)[r]),
             fa(np.asarray(pade_asyice)[r])]
        z = [fa(v) for v in (pade_sizreg_extliq, pade_sizreg_ssaliq, pade_sizreg_asyliq, pade_sizreg_extice, pade_sizreg_ssaice,
                             pade_sizreg_asyice)]
        if t[0].ndim != 3 or any(a.ndim != 3 for a in t):
            return "cloud_optics%init(): Pade coefficient arrays must be (ncoeff, nsizereg, nband)"
        ncoeff_ext, nsizereg, nbnd = t[0].shape
        ncoeff_ssa_g = t[1].shape[0]
        if t[1].shape != (ncoeff_ssa_g, nsizereg, nbnd):
            return "cloud_optics%init(): array pade_ssaliq isn't consistently sized"
        if t[2].shape != (ncoeff_ssa_g, nsizereg, nbnd):
            return "cloud_optics%init(): array pade_asyliq isn't consistently sized"
        if t[3].shape != (ncoeff_ext, nsizereg, nbnd):
            return "cloud_optics%init(): array pade_extice isn't consistently sized"
        if t[4].shape != (ncoeff_ssa_g, nsizereg, nbnd):
            return "cloud_optics%init(): array pade_ssaice isn't consistently sized"
        if t[5].shape != (ncoeff_ssa_g, nsizereg, nbnd):
            return "cloud_optics%init(): array pade_asyice isn't consistently sized"
        if any(v.shape != z[0].shape for v in z):
            return "cloud_optics%init(): one or more Pade size regime arrays are inconsistently sized"
        self.nband = self.ngpt = int(nbnd)
        self.band2gpt = np.array([[b + 1, b + 1] for b in range(self.nband)], np.int32)
        self.tables = dict(pade_extliq=t[0], pade_ssaliq=t[1], pade_asyliq=t[2], pade_extice=t[3], pade_ssaice=t[4], pade_asyice=t[5],
                           sizreg=np.stack(z))
        try:
            _lib.check(_lib.lib().rrnn_cloud_pade_create(self.ctx.h, nbnd, nsizereg, ncoeff_ext, ncoeff_ssa_g, z[0].shape[0],
                                                         *[a.ctypes.data_as(_lib.c_float_p) for a in t],
                                                         *[a.ctypes.data_as(_lib.c_float_p) for a in z], C.byref(self.h)))
        except RRNNError as e:
            return str(e)
        return ""

    def cloud_optics(self, clwp, ciwp, reliq, reice, optical_props):
        ctx = self.ctx
        a = [_dev(v, ctx) for v in (clwp, ciwp, reliq, reice)]
        ncol, nlay = a[0].shape
        if optical_props.get_ncol() != ncol or optical_props.get_nlay() != nlay:
            return "cloud optics: optical_props have wrong extents"
        if optical_props.tau.shape[-1] != self.nband:
            return "cloud optics: optical properties must be requested by band not g-points"
        two = isinstance(optical_props, ty_optical_props_2str)
        try:
            _lib.check(_lib.lib().rrnn_cloud_optics(ctx.h, self.h, ncol, nlay, *[_ptr(v) for v in a], _ptr(optical_props.tau),
                                                    _ptr(optical_props.ssa) if two else None,
                                                    _ptr(optical_props.g) if two else None))
        except RRNNError as e:
            return str(e)
        return ""

    def __del__(self):
        try:
            if self.h:
                _lib.lib().rrnn_cloud_lut_destroy(self.h)
        except Exception:
            pass


class ty_solar_var:
    """ty_solar_var (extensions/solar_variability/mo_solar_variability.F90:20-183): the mean-solar-cycle table of the facular
    and sunspot indices and its interpolation to a cycle fraction; feeds ty_gas_optics_rrtmgp%set_solar_variability.  Host-only,
    as in the reference."""

    def __init__(self):
        self.avgcyc_ind = None

    def load(self, avgcyc_ind):
        """load(avgcyc_ind), :45-69: avgcyc_ind is (nsolarterms = 2, nsolarfrac) in the reference == (nsolarfrac, 2) here."""
        a = np.ascontiguousarray(avgcyc_ind, np.float32)
        if a.ndim != 2 or a.shape[1] != 2:
            return "ty_solar_var%load: avgcyc_ind must be (nsolarfrac, 2)"
        self.avgcyc_ind = a
        return ""

    def finalize(self):
        self.avgcyc_ind = None

    def solar_var_ind_interp(self, solcycfrac):
        """-> (error_msg, mg_index, sb_index), :91-183.  As in the reference nothing is computed when no table is loaded."""
        if self.avgcyc_ind is None:
            return "", None, None
        out = np.zeros(2, np.float32)
        try:
            _lib.check(_lib.lib().rrnn_solar_var_ind_interp(self.avgcyc_ind.ctypes.data_as(_lib.c_float_p), int(self.avgcyc_ind.shape[0]),
                                                            float(solcycfrac), out[0:1].ctypes.data_as(_lib.c_float_p),
                                                            out[1:2].ctypes.data_as(_lib.c_float_p)))
        except RRNNError as e:
            return str(e), None, None
        return "", float(out[0]), float(out[1])


def load_solar_var_file(path):
    """Read extensions/solar_variability/rrtmgp-solar-var-tables.nc (classic netCDF): solar_var_avgcyc (n_solar_frac = 134,
    n_solar_terms = 2), the avgcyc_ind of ty_solar_var.load."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    a = np.array(f.variables["solar_var_avgcyc"][:], np.float32)
    f.close()
    return a if a.shape[1] == 2 else np.ascontiguousarray(a.T)


def load_cloud_lut_file(path):
    """Read a cloud-optics coefficient file (classic netCDF) -> kwargs of ty_cloud_optics.load
    (examples/all-sky/mo_load_cloud_coefficients.F90:23-110)."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    sc = lambda k: float(v[k].getValue())
    out = dict(band_lims_wvn=np.array(v["bnd_limits_wavenumber"][:], np.float32),
               radliq_lwr=sc("radliq_lwr"), radliq_upr=sc("radliq_upr"), radliq_fac=sc("radliq_fac"),
               radice_lwr=sc("radice_lwr"), radice_upr=sc("radice_upr"), radice_fac=sc("radice_fac"))
    for k in ("lut_extliq", "lut_ssaliq", "lut_asyliq", "lut_extice", "lut_ssaice", "lut_asyice"):
        out[k] = np.array(v[k][:], np.float32)
    f.close()
    return out


def _sampled_mask(name, randoms, cloud_frac, overlap_param, cloud_mask, ctx):
    torch = _torch()
    ctx = ctx or default_context()
    r, cf = _dev(randoms, ctx), _dev(cloud_frac, ctx)
    ncol, nlay, ngpt = r.shape
    if tuple(cf.shape) != (ncol, nlay):
        return "sampled_mask_max_ran: sizes of randoms(ngpt,nlay,ncol) and cloud_frac(ncol,nlay) are inconsistent"
    op = None
    if overlap_param is not None:
        op = _dev(overlap_param, ctx)
        if tuple(op.shape) != (ncol, nlay - 1):
            return "sampled_mask_max_ran: sizes of randoms(ngpt,nlay,ncol) and overlap_param(ncol,nlay-1) are inconsistent"
    if tuple(cloud_mask.shape) != (ncol, nlay, ngpt) or cloud_mask.dtype not in (torch.uint8, torch.bool):
        return "sampled_mask_max_ran: sizes of randoms(ngpt,nlay,ncol) and cloud_mask(ncol,nlay,ngpt) are inconsistent"
    if bool((cf > 1).any()) or bool((cf < 0).any()):
        return "sampled_mask_max_ran: cloud fraction values out of range [0,1]"
    if op is not None and (bool((op > 1).any()) or bool((op < -1).any())):
        return "sampled_mask_max_ran: overlap_param values out of range [-1,1]"
    try:
        _lib.check(_lib.lib().rrnn_sampled_mask(ctx.h, ngpt, nlay, ncol, _ptr(r), _ptr(cf), _ptr(op), _ptr(cloud_mask)))
    except RRNNError as e:
        return str(e)
    return ""


def sampled_mask_max_ran(randoms, cloud_frac, cloud_mask, ctx=None):
    """extensions/cloud_optics/mo_cloud_sampling.F90:107-170; randoms (ncol, nlay, ngpt), cloud_frac (ncol, nlay), cloud_mask a
    device uint8 / bool tensor (ncol, nlay, ngpt)."""
    return _sampled_mask("max_ran", randoms, cloud_frac, None, cloud_mask, ctx)


def sampled_mask_exp_ran(randoms, cloud_frac, overlap_param, cloud_mask, ctx=None):
    """extensions/cloud_optics/mo_cloud_sampling.F90:176-286; overlap_param (ncol, nlay-1).  Clear layers inside the cloudy
    range get a false mask (the reference leaves them undefined)."""
    return _sampled_mask("exp_ran", randoms, cloud_frac, overlap_param, cloud_mask, ctx)


def draw_samples(cloud_mask, clouds, clouds_sampled):
    """draw_samples (extensions/cloud_optics/mo_cloud_sampling.F90:38-101): by-band cloud properties -> sampled by g-point."""
    two, two_s = isinstance(clouds, ty_optical_props_2str), isinstance(clouds_sampled, ty_optical_props_2str)
    if two != two_s:
        return "draw_samples: by-band and sampled cloud properties need to be the same variable type"
    ncol, nlay = clouds.get_ncol(), clouds.get_nlay()
    if clouds_sampled.get_ncol() != ncol or clouds_sampled.get_nlay() != nlay:
        return "draw_samples: sampled/unsampled cloud optical properties have different ncol and/or nlay"
    if tuple(cloud_mask.shape) != (ncol, nlay, clouds_sampled.ngpt):
        return "draw_samples: cloud mask and cloud optical properties have different ncol and/or nlay"
    if clouds.tau.shape[-1] != clouds_sampled.nband:
        return "draw_samples: by-band and sampled cloud properties spectral structure is different"
    ctx = clouds_sampled.ctx
    try:
        _lib.check(_lib.lib().rrnn_draw_samples(ctx.h, clouds_sampled._kd.h, nlay, ncol, _ptr(cloud_mask), _ptr(clouds.tau),
                                                _ptr(clouds.ssa) if two else None, _ptr(clouds.g) if two else None,
                                                _ptr(clouds_sampled.tau), _ptr(clouds_sampled.ssa) if two else None,
                                                _ptr(clouds_sampled.g) if two else None))
    except RRNNError as e:
        return str(e)
    if two:
        clouds_sampled.g_is_zero = False
    return ""


def load_cloud_pade_file(path):
    """Read the Pade part of a cloud-optics coefficient file -> kwargs of ty_cloud_optics.load_pade
    (examples/all-sky/mo_load_cloud_coefficients.F90:113-200)."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    out = dict(band_lims_wvn=np.array(v["bnd_limits_wavenumber"][:], np.float32))
    for k in ("pade_extliq", "pade_ssaliq", "pade_asyliq", "pade_extice", "pade_ssaice", "pade_asyice", "pade_sizreg_extliq",
              "pade_sizreg_ssaliq", "pade_sizreg_asyliq", "pade_sizreg_extice", "pade_sizreg_ssaice", "pade_sizreg_asyice"):
        out[k] = np.array(v[k][:], np.float32)
    f.close()
    return out


def compute_heating_rate(flux_up, flux_dn, plev, heating_rate, ctx=None):
    """[K/s], extensions/mo_heating_rates.F90:26-54; arrays (ncol, nlay+1) / (ncol, nlay)."""
    ctx = ctx or default_context()
    fu, fd, pl = _dev(flux_up, ctx), _dev(flux_dn, ctx), _dev(plev, ctx)
    ncol, nlev = fu.shape
    if tuple(fd.shape) != (ncol, nlev):
        return "heating_rate: flux_dn array inconsistently sized."
    if tuple(pl.shape) != (ncol, nlev):
        return "heating_rate: plev array inconsistently sized."
    if tuple(heating_rate.shape) != (ncol, nlev - 1):
        return "heating_rate: heating_rate array inconsistently sized."
    try:
        _lib.check(_lib.lib().rrnn_heating_rate(ctx.h, ncol, nlev - 1, _ptr(fu), _ptr(fd), _ptr(pl), _ptr(heating_rate)))
    except RRNNError as e:
        return str(e)
    return ""


def calc_heating_rate(flux_up, flux_dn, pressure_hl, ctx=None):
    """[K/day], examples/rrtmgp-nn-training/rrtmgp_lw_eval_nn_rfmip.F90:624-653."""
    torch = _torch()
    ctx = ctx or default_context()
    fu, fd, pl = _dev(flux_up, ctx), _dev(flux_dn, ctx), _dev(pressure_hl, ctx)
    ncol, nlev = fu.shape
    out = torch.empty((ncol, nlev - 1), dtype=torch.float32, device=fu.device)
    _lib.check(_lib.lib().rrnn_calc_heating_rate(ctx.h, ncol, nlev - 1, _ptr(fu), _ptr(fd), _ptr(pl), _ptr(out)))
    return out


# ---- whole-path drivers (host buffers in, host buffers out) -------------------------------------------
def _models(neural_nets):
    return (vp * 2)(*[n.h for n in neural_nets][:2])


def _hp(a):
    return None if a is None else vp(a.ctypes.data)


def _flux_out(a, ncol, nlev, name):
    """A caller-supplied output array is written by the library through its raw pointer: it must be float32, C-contiguous and
    (ncol, nlay+1) -- anything else is an error, not a silent write past the buffer."""
    if a is None:
        return np.empty((ncol, nlev), np.float32)
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"] or a.shape != (ncol, nlev) \
            or not a.flags["WRITEABLE"]:
        raise ValueError(f"{name}: output array must be a writeable C-contiguous float32 array of shape ({ncol}, {nlev})")
    return a


def lw_fluxes_host(k_dist, neural_nets, play, plev, tlay, tsfc, sfc_emis, gas_desc, tlev=None, top_at_1=True,
                   n_gauss_angles=1, flux_up=None, flux_dn=None):
    """numpy in / numpy out: gas_optics(neural_nets=) -> rte_lw for all columns (rrnn_lw_fluxes_host)."""
    ctx = k_dist.ctx
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    play, plev, tlay, tlev, tsfc, sfc_emis = map(f32, (play, plev, tlay, tlev, tsfc, sfc_emis))
    ncol, nlay = play.shape
    flux_up = _flux_out(flux_up, ncol, nlay + 1, "lw_fluxes_host: flux_up")
    flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "lw_fluxes_host: flux_dn")
    gases, ngas, keep = gas_desc._to_c(ctx, host=True)
    _lib.check(_lib.lib().rrnn_lw_fluxes_host(ctx.h, k_dist._kd.h, _models(neural_nets), len(neural_nets), ncol, nlay,
                                              int(bool(top_at_1)), int(n_gauss_angles), _hp(play), _hp(plev), _hp(tlay), _hp(tlev),
                                              _hp(tsfc), _hp(sfc_emis), gases, ngas, _hp(flux_up), _hp(flux_dn)))
    return flux_up, flux_dn


def sw_fluxes_host(k_dist, neural_nets, play, plev, tlay, mu0, sfc_alb, gas_desc, tsi=None, top_at_1=True, flux_up=None,
                   flux_dn=None, flux_dn_dir=None):
    ctx = k_dist.ctx
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    play, plev, tlay, mu0, sfc_alb, tsi = map(f32, (play, plev, tlay, mu0, sfc_alb, tsi))
    ncol, nlay = play.shape
    flux_up = _flux_out(flux_up, ncol, nlay + 1, "sw_fluxes_host: flux_up")
    flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "sw_fluxes_host: flux_dn")
    flux_dn_dir = _flux_out(flux_dn_dir, ncol, nlay + 1, "sw_fluxes_host: flux_dn_dir")
    gases, ngas, keep = gas_desc._to_c(ctx, host=True)
    _lib.check(_lib.lib().rrnn_sw_fluxes_host(ctx.h, k_dist._kd.h, _models(neural_nets), ncol, nlay, int(bool(top_at_1)),
                                              _hp(play), _hp(plev), _hp(tlay), _hp(mu0), _hp(sfc_alb), _hp(tsi), gases, ngas,
                                              _hp(flux_up), _hp(flux_dn), _hp(flux_dn_dir)))
    return flux_up, flux_dn, flux_dn_dir


def lw_fluxes(k_dist, neural_nets, play, plev, tlay, tsfc, sfc_emis, gas_desc, flux_up, flux_dn, tlev=None, top_at_1=True,
              n_gauss_angles=1):
    """Device tensors in / out (rrnn_lw_fluxes)."""
    ctx = k_dist.ctx
    ncol, nlay = play.shape
    gases, ngas, keep = gas_desc._to_c(ctx)
    _lib.check(_lib.lib().rrnn_lw_fluxes(ctx.h, k_dist._kd.h, _models(neural_nets), len(neural_nets), ncol, nlay,
                                         int(bool(top_at_1)), int(n_gauss_angles), _ptr(play), _ptr(plev), _ptr(tlay), _ptr(tlev),
                                         _ptr(tsfc), _ptr(sfc_emis), gases, ngas, _ptr(flux_up), _ptr(flux_dn)))


def sw_fluxes(k_dist, neural_nets, play, plev, tlay, mu0, sfc_alb, gas_desc, flux_up, flux_dn, flux_dn_dir, tsi=None,
              top_at_1=True):
    ctx = k_dist.ctx
    ncol, nlay = play.shape
    gases, ngas, keep = gas_desc._to_c(ctx)
    _lib.check(_lib.lib().rrnn_sw_fluxes(ctx.h, k_dist._kd.h, _models(neural_nets), ncol, nlay, int(bool(top_at_1)), _ptr(play),
                                         _ptr(plev), _ptr(tlay), _ptr(mu0), _ptr(sfc_alb), _ptr(tsi), gases, ngas,
                                         _ptr(flux_up), _ptr(flux_dn), _ptr(flux_dn_dir)))


def _cloud_args(cl, conv):
    """clwp, ciwp, reliq, reice of a dict with the keys of synth.make_clouds / drivers.allsky_clouds"""
    return [conv(cl[k]) for k in ("lwp", "iwp", "rel", "rei")]


def lw_fluxes_allsky(k_dist, neural_nets, cloud_optics, play, plev, tlay, tsfc, sfc_emis, gas_desc, clouds, flux_up, flux_dn, tlev=None,
                     top_at_1=True, n_gauss_angles=1):
    """Device tensors in / out: cloud_optics + gas_optics(neural_nets=) + increment + rte_lw for all columns (rrnn_lw_fluxes_allsky);
    cloud_optics is a loaded ty_cloud_optics, clouds a dict lwp / iwp / rel / rei of (ncol, nlay) tensors."""
    ctx = k_dist.ctx
    ncol, nlay = play.shape
    gases, ngas, keep = gas_desc._to_c(ctx)
    _lib.check(_lib.lib().rrnn_lw_fluxes_allsky(ctx.h, k_dist._kd.h, _models(neural_nets), len(neural_nets), cloud_optics.h, ncol, nlay,
                                                int(bool(top_at_1)), int(n_gauss_angles), _ptr(play), _ptr(plev), _ptr(tlay), _ptr(tlev),
                                                _ptr(tsfc), _ptr(sfc_emis), gases, ngas, *_cloud_args(clouds, _ptr), _ptr(flux_up),
                                                _ptr(flux_dn)))


def sw_fluxes_allsky(k_dist, neural_nets, cloud_optics, play, plev, tlay, mu0, sfc_alb, gas_desc, clouds, flux_up, flux_dn, flux_dn_dir,
                     tsi=None, top_at_1=True):
    ctx = k_dist.ctx
    ncol, nlay = play.shape
    gases, ngas, keep = gas_desc._to_c(ctx)
    _lib.check(_lib.lib().rrnn_sw_fluxes_allsky(ctx.h, k_dist._kd.h, _models(neural_nets), cloud_optics.h, ncol, nlay, int(bool(top_at_1)),
                                                _ptr(play), _ptr(plev), _ptr(tlay), _ptr(mu0), _ptr(sfc_alb), _ptr(tsi), gases, ngas,
                                                *_cloud_args(clouds, _ptr), _ptr(flux_up), _ptr(flux_dn), _ptr(flux_dn_dir)))


def lw_fluxes_allsky_host(k_dist, neural_nets, cloud_optics, play, plev, tlay, tsfc, sfc_emis, gas_desc, clouds, tlev=None, top_at_1=True,
                          n_gauss_angles=1, flux_up=None, flux_dn=None):
    """numpy in / numpy out (rrnn_lw_fluxes_allsky_host)."""
    ctx = k_dist.ctx
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    play, plev, tlay, tlev, tsfc, sfc_emis = map(f32, (play, plev, tlay, tlev, tsfc, sfc_emis))
    cl = _cloud_args(clouds, f32)
    ncol, nlay = play.shape
    flux_up = _flux_out(flux_up, ncol, nlay + 1, "lw_fluxes_allsky_host: flux_up")
    flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "lw_fluxes_allsky_host: flux_dn")
    gases, ngas, keep = gas_desc._to_c(ctx, host=True)
    _lib.check(_lib.lib().rrnn_lw_fluxes_allsky_host(ctx.h, k_dist._kd.h, _models(neural_nets), len(neural_nets), cloud_optics.h, ncol, nlay,
                                                     int(bool(top_at_1)), int(n_gauss_angles), _hp(play), _hp(plev), _hp(tlay), _hp(tlev),
                                                     _hp(tsfc), _hp(sfc_emis), gases, ngas, *[_hp(a) for a in cl], _hp(flux_up), _hp(flux_dn)))
    return flux_up, flux_dn


def sw_fluxes_allsky_host(k_dist, neural_nets, cloud_optics, play, plev, tlay, mu0, sfc_alb, gas_desc, clouds, tsi=None, top_at_1=True,
                          flux_up=None, flux_dn=None, flux_dn_dir=None):
    ctx = k_dist.ctx
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    play, plev, tlay, mu0, sfc_alb, tsi = map(f32, (play, plev, tlay, mu0, sfc_alb, tsi))
    cl = _cloud_args(clouds, f32)
    ncol, nlay = play.shape
    flux_up = _flux_out(flux_up, ncol, nlay + 1, "sw_fluxes_allsky_host: flux_up")
    flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "sw_fluxes_allsky_host: flux_dn")
    flux_dn_dir = _flux_out(flux_dn_dir, ncol, nlay + 1, "sw_fluxes_allsky_host: flux_dn_dir")
    gases, ngas, keep = gas_desc._to_c(ctx, host=True)
    _lib.check(_lib.lib().rrnn_sw_fluxes_allsky_host(ctx.h, k_dist._kd.h, _models(neural_nets), cloud_optics.h, ncol, nlay,
                                                     int(bool(top_at_1)), _hp(play), _hp(plev), _hp(tlay), _hp(mu0), _hp(sfc_alb), _hp(tsi),
                                                     gases, ngas, *[_hp(a) for a in cl], _hp(flux_up), _hp(flux_dn), _hp(flux_dn_dir)))
    return flux_up, flux_dn, flux_dn_dir


class MultiDevice:
    """One process, N devices (rrnn_multi_*): contexts, spectral tables and networks replicated per device, the columns of a
    call cut into contiguous shards, every shard's host pipeline on its own thread, fluxes written into the caller's host arrays."""

    def __init__(self, devices=None, ndev=None):
        self.h = vp()
        if devices is None:
            n = int(ndev or _lib.lib().rrnn_device_count())
            _lib.check(_lib.lib().rrnn_multi_create(n, None, C.byref(self.h)))
        else:
            d = np.ascontiguousarray(devices, np.int32)
            _lib.check(_lib.lib().rrnn_multi_create(len(d), d.ctypes.data_as(_lib.c_int_p), C.byref(self.h)))
        self.ndev = int(_lib.lib().rrnn_multi_ndev(self.h))

    def set_flag(self, name, value):
        _lib.check(_lib.lib().rrnn_multi_set_flag(self.h, name.encode(), int(value)))

    def load_netcdf(self, filename):
        i = C.c_int(-1)
        _lib.check(_lib.lib().rrnn_multi_model_load_netcdf(self.h, str(filename).encode(), C.byref(i)))
        return int(i.value)

    def load_kdist(self, kd):
        bl = np.ascontiguousarray(kd["band_lims_gpt"], np.int32)
        tot, sol = kd.get("totplnk"), kd.get("solar_source")
        tot_a = None if tot is None else np.ascontiguousarray(tot, np.float32)
        sol_a = None if sol is None else np.ascontiguousarray(sol, np.float32)
        fp = lambda a: None if a is None else a.ctypes.data_as(_lib.c_float_p)
        i = C.c_int(-1)
        _lib.check(_lib.lib().rrnn_multi_kdist_create(self.h, int(kd["nbnd"]), int(kd["ngpt"]), bl.ctypes.data_as(_lib.c_int_p),
                                                      0 if tot_a is None else tot_a.shape[1], fp(tot_a), float(kd.get("temp_ref_min", 0.0)),
                                                      float(kd.get("totplnk_delta", 1.0)), fp(sol_a), C.byref(i)))
        return int(i.value)

    def lw_fluxes_host(self, kdist_id, model_ids, play, plev, tlay, tsfc, sfc_emis, gas_desc, tlev=None, top_at_1=True,
                       n_gauss_angles=1, flux_up=None, flux_dn=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        play, plev, tlay, tlev, tsfc, sfc_emis = map(f32, (play, plev, tlay, tlev, tsfc, sfc_emis))
        ncol, nlay = play.shape
        flux_up = _flux_out(flux_up, ncol, nlay + 1, "lw_fluxes_host: flux_up")
        flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "lw_fluxes_host: flux_dn")
        gases, ngas, keep = gas_desc._to_c(None, host=True)
        ids = np.ascontiguousarray(model_ids, np.int32)
        _lib.check(_lib.lib().rrnn_multi_lw_fluxes_host(self.h, int(kdist_id), ids.ctypes.data_as(_lib.c_int_p), len(ids), ncol, nlay,
                                                        int(bool(top_at_1)), int(n_gauss_angles), _hp(play), _hp(plev), _hp(tlay), _hp(tlev),
                                                        _hp(tsfc), _hp(sfc_emis), gases, ngas, _hp(flux_up), _hp(flux_dn)))
        return flux_up, flux_dn

    def sw_fluxes_host(self, kdist_id, model_ids, play, plev, tlay, mu0, sfc_alb, gas_desc, tsi=None, top_at_1=True, flux_up=None,
                       flux_dn=None, flux_dn_dir=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        play, plev, tlay, mu0, sfc_alb, tsi = map(f32, (play, plev, tlay, mu0, sfc_alb, tsi))
        ncol, nlay = play.shape
        flux_up = _flux_out(flux_up, ncol, nlay + 1, "sw_fluxes_host: flux_up")
        flux_dn = _flux_out(flux_dn, ncol, nlay + 1, "sw_fluxes_host: flux_dn")
        flux_dn_dir = _flux_out(flux_dn_dir, ncol, nlay + 1, "sw_fluxes_host: flux_dn_dir")
        gases, ngas, keep = gas_desc._to_c(None, host=True)
        ids = np.ascontiguousarray(model_ids, np.int32)
        _lib.check(_lib.lib().rrnn_multi_sw_fluxes_host(self.h, int(kdist_id), ids.ctypes.data_as(_lib.c_int_p), ncol, nlay,
                                                        int(bool(top_at_1)), _hp(play), _hp(plev), _hp(tlay), _hp(mu0), _hp(sfc_alb),
                                                        _hp(tsi), gases, ngas, _hp(flux_up), _hp(flux_dn), _hp(flux_dn_dir)))
        return flux_up, flux_dn, flux_dn_dir

    def close(self):
        if self.h:
            _lib.lib().rrnn_multi_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
