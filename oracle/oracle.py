"""ctypes front end of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Parity pinned IN PART (see oracle.c header): the gas-optics pre- and post-processing against the reference's own Python
(tests/golden/ref_python_golden.npz); "PARITY UNPINNED" for the solvers -- the reference Fortran cannot be built in this image.

The orchestration functions at the bottom restate the reference's callers:
  gas_optics_lw  <- rrtmgp/mo_gas_optics_rrtmgp.F90:239-428 (NN branch :368-411)
  gas_optics_sw  <- rrtmgp/mo_gas_optics_rrtmgp.F90:433-602 (NN branch :529-573, toa :594-599)
  rte_lw         <- rte/mo_rte_lw.F90:60-424
  rte_sw         <- rte/mo_rte_sw.F90:48-266
Arrays are numpy, C-order, g-point fastest: tau[ncol, nlay, ngpt].

Every function takes `fast`, which selects the build of oracle.c:
  False  strict IEEE fp32, no contraction      -> THE parity checker (the reference's working precision)
  True   -O3 / AVX2+FMA / OpenMP fp32          -> reported CPU baseline
  "f64"  the same algorithm in double          -> rounding-free yardstick: tells implementation differences
                                                  from the fp32 rounding noise of the reference arithmetic itself
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ACT = {"linear": 0, "softsign": 1, "relu": 2, "sigmoid": 3, "hard_sigmoid": 4}

# rte/mo_rte_lw.F90:113-125
GAUSS_DS = np.array([[1.66, 0, 0, 0], [1.18350343, 2.81649655, 0, 0], [1.09719858, 1.69338507, 4.70941630, 0],
                     [1.06056257, 1.38282560, 2.40148179, 7.15513024]], dtype=np.float32)
GAUSS_WTS = np.array([[0.5, 0, 0, 0], [0.3180413817, 0.1819586183, 0, 0],
                      [0.2009319137, 0.2292411064, 0.0698269799, 0],
                      [0.1355069134, 0.2034645680, 0.1298475476, 0.0311809710]], dtype=np.float32)

_NAMES = {"strict": "liboracle.so", "fast": "liboracle_fast.so", "f64": "liboracle_f64.so"}


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    stale = any(not os.path.exists(os.path.join(_HERE, "_build", n)) or
                os.path.getmtime(os.path.join(_HERE, "_build", n)) < os.path.getmtime(src) for n in _NAMES.values())
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])


_libs = {}


def _key(fast):
    return "f64" if fast == "f64" else ("fast" if fast else "strict")


def lib(fast=False):
    key = _key(fast)
    if key not in _libs:
        build()
        _libs[key] = C.CDLL(os.path.join(_HERE, "_build", _NAMES[key]))
    return _libs[key]


def _dt(fast):
    return np.float64 if fast == "f64" else np.float32


def _sc(v, fast):
    return C.c_double(float(v)) if fast == "f64" else C.c_float(float(v))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _a(x, fast):
    """contiguous array of the build's real kind"""
    return np.ascontiguousarray(x, dtype=_dt(fast))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Net:
    """Weights exactly as in the file: W[n] row-major (n_in, n_out)."""

    def __init__(self, model):
        self.dims = [int(d) for d in model["dims"]]
        self.nlayers = len(self.dims) - 1
        self.wpack = np.concatenate([np.ascontiguousarray(w, np.float32).ravel() for w in model["W"]])
        self.bpack = np.concatenate([np.asarray(b, np.float32).ravel() for b in model["b"]])
        self.act = np.array([ACT[a] for a in model["activations"]], dtype=np.int32)
        self.dims_a = np.array(self.dims, dtype=np.int32)
        self.ymean = None if model.get("ymean") is None else np.asarray(model["ymean"], np.float32)
        self.ystd = None if model.get("ystd") is None else np.asarray(model["ystd"], np.float32)
        self.xmin = np.asarray(model["xmin"], np.float32)
        self.xmax = np.asarray(model["xmax"], np.float32)
        self.input_names = list(model["input_names"])
        self._cache = {}

    def arr(self, name, fast):
        k = (name, _key(fast) == "f64")
        if k not in self._cache:
            self._cache[k] = _a(getattr(self, name), fast)
        return self._cache[k]

    def _args(self, fast):
        return (C.c_int(self.nlayers), _ip(self.dims_a), _p(self.arr("wpack", fast)), _p(self.arr("bpack", fast)), _ip(self.act))


def get_col_dry(vmr_h2o, plev, fast=False):
    ncol, nlay = np.shape(vmr_h2o)
    h, p = _a(vmr_h2o, fast), _a(plev, fast)
    out = np.empty((ncol, nlay), _dt(fast))
    lib(fast).orc_get_col_dry(ncol, nlay, _p(h), _p(p), _p(out))
    return out


def interp_tlev(play, plev, tlay, fast=False):
    ncol, nlay = play.shape
    a, b, c = _a(play, fast), _a(plev, fast), _a(tlay, fast)
    out = np.empty((ncol, nlay + 1), _dt(fast))
    lib(fast).orc_interp_tlev(ncol, nlay, _p(a), _p(b), _p(c), _p(out))
    return out


def compute_nn_inputs(net, play, tlay, gases, fast=False):
    """gases: dict name -> scalar | (nlay,) | (ncol,nlay) array (ty_gas_concs semantics,
    rrtmgp/mo_gas_concentrations.F90:50-88).  Matching is by name (compute_nn_inputs :708-760)."""
    ncol, nlay = play.shape
    nx = net.dims[0]
    ptrs = (C.c_void_p * nx)()
    modes = np.full(nx, -1, np.int32)
    keep = []
    for i, name in enumerate(net.input_names):
        if i < 2:
            continue
        if name not in gases:
            if i < 4:
                raise ValueError("h2o and o3 are required")
            continue
        v = np.asarray(gases[name], np.float32)
        if i < 4:
            v = np.broadcast_to(v, (ncol, nlay)); modes[i] = 2
        elif v.ndim == 0 or v.size == 1:
            v = v.reshape(1); modes[i] = 0
        elif v.ndim == 1:
            assert v.shape == (nlay,); modes[i] = 1
        else:
            assert v.shape == (ncol, nlay); modes[i] = 2
        v = _a(v, fast)
        keep.append(v)
        ptrs[i] = v.ctypes.data
    p, t = _a(play, fast), _a(tlay, fast)
    out = np.empty((ncol, nlay, nx), _dt(fast))
    lib(fast).orc_compute_nn_inputs(ncol, nlay, nx, _p(p), _p(t), ptrs, _ip(modes), _p(net.arr("xmin", fast)),
                                    _p(net.arr("xmax", fast)), _p(out))
    return out


def output_sgemm_tau(net, x, coldry, output2=None, fast=False):
    x = _a(np.reshape(x, (-1, net.dims[0])), fast)
    nb = x.shape[0]
    cd = _a(np.reshape(coldry, -1), fast)
    out = np.empty((nb, net.dims[-1]), _dt(fast))
    o2p = None
    if output2 is not None:
        assert output2.dtype == _dt(fast) and output2.flags.c_contiguous
        o2p = _p(output2)
    lib(fast).orc_output_sgemm_tau(*net._args(fast), _p(net.arr("ymean", fast)), _p(net.arr("ystd", fast)), nb, _p(x), _p(cd),
                                   _p(out), o2p)
    return out


def output_sgemm_pfrac(net, x, fast=False):
    x = _a(np.reshape(x, (-1, net.dims[0])), fast)
    nb = x.shape[0]
    out = np.empty((nb, net.dims[-1]), _dt(fast))
    lib(fast).orc_output_sgemm_pfrac(*net._args(fast), nb, _p(x), _p(out))
    return out


def output_sgemm_lw(net, x, fast=False):
    x = _a(np.reshape(x, (-1, net.dims[0])), fast)
    nb = x.shape[0]
    out = np.empty((nb, net.dims[-1]), _dt(fast))
    lib(fast).orc_output_sgemm_lw(*net._args(fast), nb, _p(x), _p(out))
    return out


def predict_nn_lw(nets, nn_inputs, col_dry, ngpt, fast=False):
    """predict_nn_lw_blas_sp: rrtmgp/kernels/mo_gas_optics_kernels.F90:690-774 -> (tau, pfrac)."""
    ncol, nlay, nx = nn_inputs.shape
    if len(nets) == 2:
        tau = output_sgemm_tau(nets[0], nn_inputs, col_dry, fast=fast)
        pfrac = output_sgemm_pfrac(nets[1], nn_inputs, fast=fast)
    else:
        both = output_sgemm_lw(nets[0], nn_inputs, fast=fast)
        tau = np.empty((ncol * nlay, ngpt), _dt(fast))
        pfrac = np.empty((ncol * nlay, ngpt), _dt(fast))
        cd = _a(np.reshape(col_dry, -1), fast)
        ym = _a(nets[0].ymean[:ngpt], fast); ys = _a(nets[0].ystd[:ngpt], fast)
        lib(fast).orc_split_both(ncol, nlay, ngpt, _p(both), _p(ym), _p(ys), _p(cd), _p(tau), _p(pfrac))
    return tau.reshape(ncol, nlay, ngpt), pfrac.reshape(ncol, nlay, ngpt)


def predict_nn_sw(nets, nn_inputs, col_dry, fast=False):
    """predict_nn_sw_blas_sp: rrtmgp/kernels/mo_gas_optics_kernels.F90:869-953 -> (tau_tot, ssa)."""
    ncol, nlay, nx = nn_inputs.shape
    tau = output_sgemm_tau(nets[0], nn_inputs, col_dry, fast=fast)               # tau_abs
    ssa = output_sgemm_tau(nets[1], nn_inputs, col_dry, output2=tau, fast=fast)  # tau -> tau_tot, returns ssa
    ngpt = tau.shape[-1]
    return tau.reshape(ncol, nlay, ngpt), ssa.reshape(ncol, nlay, ngpt)


def planck_source_nn(kd, tlay, tlev, tsfc, sfc_lay, pfrac, fast=False):
    """compute_Planck_source_nn; returns (sfc_source, sfc_source_Jac, lay_source, lev_source)."""
    ncol, nlay, ngpt = pfrac.shape
    dt = _dt(fast)
    lay = np.array(pfrac, dtype=dt, order="C", copy=True)
    lev = np.empty((ncol, nlay + 1, ngpt), dt)
    sfc = np.empty((ncol, ngpt), dt)
    jac = np.empty((ncol, ngpt), dt)
    a, b, c = _a(tlay, fast), _a(tlev, fast), _a(tsfc, fast)
    bl = np.ascontiguousarray(kd["band_lims_gpt"], np.int32)
    tp = _a(kd["totplnk"], fast)
    nbnd, ntemp = tp.shape
    lib(fast).orc_planck_source_nn(ncol, nlay, nbnd, ngpt, ntemp, _p(a), _p(b), _p(c), int(sfc_lay), _ip(bl),
                                   _sc(kd["temp_ref_min"], fast), _sc(kd["totplnk_delta"], fast), _p(tp), _p(sfc), _p(jac),
                                   _p(lay), _p(lev))
    return sfc, jac, lay, lev


def gas_optics_lw(kd, nets, play, plev, tlay, tsfc, gases, tlev=None, fast=False):
    """ty_gas_optics_rrtmgp%gas_optics (LW, neural_nets present)."""
    ncol, nlay = play.shape
    if tlev is None:
        tlev = interp_tlev(play, plev, tlay, fast)
    h2o = np.broadcast_to(np.asarray(gases["h2o"], np.float32), (ncol, nlay))
    col_dry = get_col_dry(h2o, plev, fast)
    x = compute_nn_inputs(nets[0], play, tlay, gases, fast)
    tau, pfrac = predict_nn_lw(nets, x, col_dry, kd["ngpt"], fast)
    sfc_lay = 1 if play[0, 0] > play[0, nlay - 1] else nlay
    sfc, jac, lay, lev = planck_source_nn(kd, tlay, tlev, tsfc, sfc_lay, pfrac, fast)
    return dict(tau=tau, lay_source=lay, lev_source=lev, sfc_source=sfc, sfc_source_Jac=jac, col_dry=col_dry,
                nn_inputs=x, tlev=tlev, pfrac=pfrac)


def gas_optics_sw(kd, nets, play, plev, tlay, gases, fast=False):
    """ty_gas_optics_rrtmgp%gas_optics (SW, neural_nets present, 2str)."""
    ncol, nlay = play.shape
    h2o = np.broadcast_to(np.asarray(gases["h2o"], np.float32), (ncol, nlay))
    col_dry = get_col_dry(h2o, plev, fast)
    x = compute_nn_inputs(nets[0], play, tlay, gases, fast)
    tau, ssa = predict_nn_sw(nets, x, col_dry, fast)
    g = np.zeros_like(tau)
    toa = np.ascontiguousarray(np.broadcast_to(np.asarray(kd["solar_source"], _dt(fast)), (ncol, kd["ngpt"])))
    return dict(tau=tau, ssa=ssa, g=g, toa_src=toa, col_dry=col_dry, nn_inputs=x)


def expand(band_lims_gpt, ngpt, arr, fast=False):
    ncol, nband = np.shape(arr)
    a = _a(arr, fast); bl = np.ascontiguousarray(band_lims_gpt, np.int32)
    out = np.empty((ncol, ngpt), _dt(fast))
    lib(fast).orc_expand(nband, ngpt, ncol, _ip(bl), _p(a), _p(out))
    return out


def lw_solver_noscat_GaussQuad(top_at_1, nmus, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source,
                               inc_flux=None, fast=False):
    ncol, nlay, ngpt = tau.shape
    dt = _dt(fast)
    if inc_flux is None:
        inc_flux = np.zeros((ncol, ngpt), dt)
    Ds = _a(GAUSS_DS[nmus - 1, :nmus], fast); wts = _a(GAUSS_WTS[nmus - 1, :nmus], fast)
    a = [_a(v, fast) for v in (inc_flux, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source)]
    up = np.empty((ncol, nlay + 1), dt); dn = np.empty((ncol, nlay + 1), dt)
    lib(fast).orc_lw_solver_noscat_GaussQuad(ngpt, nlay, ncol, int(bool(top_at_1)), nmus, _p(Ds), _p(wts),
                                             *[_p(v) for v in a], _p(up), _p(dn))
    return up, dn


def rte_lw(kd, top_at_1, tau, lay_source, lev_source, sfc_source, sfc_emis, n_gauss_angles=1, inc_flux=None,
           fast=False):
    """rte_lw for ty_optical_props_1scl: expand emissivity by band, zero incident flux, GaussQuad solver."""
    ngpt = tau.shape[-1]
    emis_gpt = expand(kd["band_lims_gpt"], ngpt, sfc_emis, fast)
    return lw_solver_noscat_GaussQuad(top_at_1, n_gauss_angles, tau, lay_source, lev_source, emis_gpt, sfc_source,
                                      inc_flux, fast)


def lw_solver_noscat_GaussQuad_ext(top_at_1, nmus, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source, inc_flux=None,
                                   ssa=None, g=None, lw_Ds=None, sfc_source_Jac=None, want_gpt=False, fast=False):
    """Everything rte_lw can ask of lw_solver_noscat_GaussQuad (rte/mo_rte_lw.F90:324-384): re-scaled scattering (ssa, g),
    per-g-point secants lw_Ds (ncol, ngpt; one angle), the surface-temperature Jacobian, g-point fluxes.
    Returns a dict with flux_up, flux_dn [, flux_up_Jac][, gpt_flux_up, gpt_flux_dn]."""
    ncol, nlay, ngpt = tau.shape
    dt = _dt(fast)
    if inc_flux is None:
        inc_flux = np.zeros((ncol, ngpt), dt)
    Ds = _a(GAUSS_DS[nmus - 1, :nmus], fast); wts = _a(GAUSS_WTS[nmus - 1, :nmus], fast)
    opt = lambda v: None if v is None else _a(v, fast)
    ssa_a, g_a, dsg, sj = opt(ssa), opt(g), opt(lw_Ds), opt(sfc_source_Jac)
    inc, tau_a, lay, lev, emis, ssrc = [_a(v, fast) for v in (inc_flux, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source)]
    up = np.empty((ncol, nlay + 1), dt); dn = np.empty_like(up)
    jac = np.empty_like(up) if sj is not None else None
    gup = np.empty((ncol, nlay + 1, ngpt), dt) if want_gpt else None
    gdn = np.empty_like(gup) if want_gpt else None
    pn = lambda v: None if v is None else _p(v)
    lib(fast).orc_lw_solver_noscat_GaussQuad_ext(ngpt, nlay, ncol, int(bool(top_at_1)), nmus, _p(Ds), _p(wts), pn(dsg), _p(inc),
                                                 _p(tau_a), pn(ssa_a), pn(g_a), _p(lay), _p(lev), _p(emis), _p(ssrc), pn(sj),
                                                 _p(up), _p(dn), pn(jac), pn(gup), pn(gdn))
    out = dict(flux_up=up, flux_dn=dn)
    if jac is not None:
        out["flux_up_Jac"] = jac
    if want_gpt:
        out["gpt_flux_up"], out["gpt_flux_dn"] = gup, gdn
    return out


def lw_solver_2stream(top_at_1, tau, ssa, g, lev_source, sfc_emis_gpt, sfc_source, inc_flux=None, want_gpt=False, fast=False):
    """lw_solver_2stream (rte/kernels/mo_rte_solver_kernels.F90:426-486): flux_up, flux_dn [, gpt_flux_up, gpt_flux_dn]."""
    ncol, nlay, ngpt = tau.shape
    dt = _dt(fast)
    if inc_flux is None:
        inc_flux = np.zeros((ncol, ngpt), dt)
    a = [_a(v, fast) for v in (inc_flux, tau, ssa, g, lev_source, sfc_emis_gpt, sfc_source)]
    up = np.empty((ncol, nlay + 1), dt); dn = np.empty_like(up)
    gu = np.empty((ncol, nlay + 1, ngpt), dt) if want_gpt else None
    gd = np.empty_like(gu) if want_gpt else None
    lib(fast).orc_lw_solver_2stream(ngpt, nlay, ncol, int(bool(top_at_1)), *[_p(v) for v in a], _p(up), _p(dn),
                                    None if gu is None else _p(gu), None if gd is None else _p(gd))
    return (up, dn, gu, gd) if want_gpt else (up, dn)


def sw_solver_2stream(top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif, fast=False):
    ncol, nlay, ngpt = tau.shape
    dt = _dt(fast)
    a = [_a(v, fast) for v in (inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif)]
    up = np.empty((ncol, nlay + 1), dt); dn = np.empty_like(up); dr = np.empty_like(up)
    lib(fast).orc_sw_solver_2stream(ngpt, nlay, ncol, int(bool(top_at_1)), *[_p(v) for v in a], _p(up), _p(dn), _p(dr))
    return up, dn, dr


def sw_solver_2stream_gpt(top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif, fast=False):
    """sw_solver_2stream with the optional g-point fluxes (mo_rte_solver_kernels.F90:541-692, save_gpt_flux): returns
    flux_up, flux_dn, flux_dir, gpt_flux_up, gpt_flux_dn (total), gpt_flux_dn_dir."""
    ncol, nlay, ngpt = tau.shape
    dt = _dt(fast)
    a = [_a(v, fast) for v in (inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif)]
    up = np.empty((ncol, nlay + 1), dt); dn = np.empty_like(up); dr = np.empty_like(up)
    gu = np.empty((ncol, nlay + 1, ngpt), dt); gd = np.empty_like(gu); gr = np.empty_like(gu)
    lib(fast).orc_sw_solver_2stream_gpt(ngpt, nlay, ncol, int(bool(top_at_1)), *[_p(v) for v in a], _p(up), _p(dn), _p(dr),
                                        _p(gu), _p(gd), _p(gr))
    return up, dn, dr, gu, gd, gr


def rte_sw(top_at_1, mu0, inc_flux, alb_dir_gpt, alb_dif_gpt, tau, ssa, g, inc_flux_dif=None, fast=False):
    if inc_flux_dif is None:
        inc_flux_dif = np.zeros_like(inc_flux)
    return sw_solver_2stream(top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir_gpt, alb_dif_gpt, fast)


def cloud_optics_lut(co, clwp, ciwp, reliq, reice, two_stream, fast=False):
    """co: tables dict of rte_rrtmgp_nn_b200.api.ty_cloud_optics (tables [nbnd][nsteps], ice roughness chosen)."""
    ncol, nlay = clwp.shape
    dt = _dt(fast)
    nbnd = co["extliq"].shape[0]
    a = [_a(v, fast) for v in (clwp, ciwp, reliq, reice)]
    t = [_a(co[k], fast) for k in ("extliq", "ssaliq", "asyliq", "extice", "ssaice", "asyice")]
    tau = np.empty((ncol, nlay, nbnd), dt); ssa = np.empty_like(tau); g = np.empty_like(tau)
    lib(fast).orc_cloud_optics_lut(ncol, nlay, nbnd, *[_p(v) for v in a],
                                   int(co["liq_nsteps"]), _sc(co["liq_step_size"], fast), _sc(co["radliq_lwr"], fast),
                                   _p(t[0]), _p(t[1]), _p(t[2]),
                                   int(co["ice_nsteps"]), _sc(co["ice_step_size"], fast), _sc(co["radice_lwr"], fast),
                                   _p(t[3]), _p(t[4]), _p(t[5]), int(two_stream), _p(tau), _p(ssa), _p(g))
    if two_stream:
        return tau, ssa, g
    return tau


def cloud_optics_pade(co, clwp, ciwp, reliq, reice, two_stream, fast=False):
    """co: dict with pade_{ext,ssa,asy}{liq,ice} (ncoeff, 3, nbnd; ice for one roughness) and sizreg (6, 4)
    (extensions/cloud_optics/mo_cloud_optics.F90:476-528, 650-781)."""
    ncol, nlay = clwp.shape
    dt = _dt(fast)
    nbnd = co["pade_extliq"].shape[-1]
    a = [_a(v, fast) for v in (clwp, ciwp, reliq, reice)]
    t = [_a(co[k], fast) for k in ("pade_extliq", "pade_ssaliq", "pade_asyliq", "pade_extice", "pade_ssaice", "pade_asyice")]
    sz = _a(co["sizreg"], fast)
    tau = np.empty((ncol, nlay, nbnd), dt); ssa = np.empty_like(tau); g = np.empty_like(tau)
    lib(fast).orc_cloud_optics_pade(ncol, nlay, nbnd, *[_p(v) for v in a], *[_p(v) for v in t], _p(sz), int(two_stream),
                                    _p(tau), _p(ssa), _p(g))
    if two_stream:
        return tau, ssa, g
    return tau


def sampled_mask(randoms, cloud_frac, overlap_param=None, fast=False):
    """sampled_mask_max_ran (overlap_param None) / sampled_mask_exp_ran, extensions/cloud_optics/mo_cloud_sampling.F90:107-286;
    randoms (ncol, nlay, ngpt), cloud_frac (ncol, nlay), overlap_param (ncol, nlay-1) -> bool mask (ncol, nlay, ngpt)."""
    ncol, nlay, ngpt = randoms.shape
    r = _a(randoms, fast); cf = _a(cloud_frac, fast)
    op = None if overlap_param is None else _a(overlap_param, fast)
    m = np.empty((ncol, nlay, ngpt), np.uint8)
    lib(fast).orc_sampled_mask(ngpt, nlay, ncol, _p(r), _p(cf), None if op is None else _p(op), _p(m))
    return m.astype(bool)


def draw_samples(cloud_mask, gpt_lims, *fields, fast=False):
    """draw_samples / apply_cloud_mask (:38-101, 292-308): by-band fields (ncol, nlay, nbnd) -> sampled (ncol, nlay, ngpt)."""
    ncol, nlay, ngpt = cloud_mask.shape
    m = np.ascontiguousarray(cloud_mask, np.uint8); gl = np.ascontiguousarray(gpt_lims, np.int32)
    out = []
    for f in fields:
        a = _a(f, fast); o = np.zeros((ncol, nlay, ngpt), _dt(fast))
        lib(fast).orc_apply_cloud_mask(ngpt, nlay, ncol, a.shape[-1], _ip(gl), _p(m), _p(a), _p(o))
        out.append(o)
    return out


def delta_scale_2str(tau, ssa, g, fast=False):
    dt = _dt(fast)
    tau = np.array(tau, dt, copy=True); ssa = np.array(ssa, dt, copy=True); g = np.array(g, dt, copy=True)
    lib(fast).orc_delta_scale_2str(C.c_size_t(tau.size), _p(tau), _p(ssa), _p(g))
    return tau, ssa, g


def inc_1scalar_by_1scalar_bybnd(tau1, tau2, gpt_lims, fast=False):
    ncol, nlay, ngpt = tau1.shape
    tau1 = np.array(tau1, _dt(fast), copy=True)
    t2 = _a(tau2, fast); gl = np.ascontiguousarray(gpt_lims, np.int32)
    lib(fast).orc_inc_1scalar_by_1scalar_bybnd(ngpt, nlay, ncol, _p(tau1), _p(t2), t2.shape[-1], _ip(gl))
    return tau1


def inc_2stream_by_2stream_bybnd(tau1, ssa1, g1, tau2, ssa2, g2, gpt_lims, fast=False):
    ncol, nlay, ngpt = tau1.shape
    dt = _dt(fast)
    tau1 = np.array(tau1, dt, copy=True); ssa1 = np.array(ssa1, dt, copy=True); g1 = np.array(g1, dt, copy=True)
    a = [_a(v, fast) for v in (tau2, ssa2, g2)]; gl = np.ascontiguousarray(gpt_lims, np.int32)
    lib(fast).orc_inc_2stream_by_2stream_bybnd(ngpt, nlay, ncol, _p(tau1), _p(ssa1), _p(g1), _p(a[0]), _p(a[1]), _p(a[2]),
                                               a[0].shape[-1], _ip(gl))
    return tau1, ssa1, g1


def heating_rate(flux_up, flux_dn, plev, fast=False):
    ncol, nlev = flux_up.shape
    a = [_a(v, fast) for v in (flux_up, flux_dn, plev)]
    out = np.empty((ncol, nlev - 1), _dt(fast))
    lib(fast).orc_heating_rate(ncol, nlev - 1, _p(a[0]), _p(a[1]), _p(a[2]), _p(out))
    return out


def calc_heating_rate(flux_up, flux_dn, plev, fast=False):
    ncol, nlev = flux_up.shape
    a = [_a(v, fast) for v in (flux_up, flux_dn, plev)]
    out = np.empty((ncol, nlev - 1), _dt(fast))
    lib(fast).orc_calc_heating_rate(ncol, nlev - 1, _p(a[0]), _p(a[1]), _p(a[2]), _p(out))
    return out


def sum_byband(gpt_flux, band_lims, fast=False):
    """sum_byband (extensions/mo_fluxes_byband_kernels.F90:33-51): (ncol, nlev, ngpt) -> (ncol, nlev, nbnd)."""
    ncol, nlev, ngpt = np.shape(gpt_flux)
    f = _a(gpt_flux, fast); bl = np.ascontiguousarray(band_lims, np.int32)
    out = np.empty((ncol, nlev, len(bl)), _dt(fast))
    lib(fast).orc_sum_byband(ncol, nlev, ngpt, len(bl), _ip(bl), _p(f), _p(out))
    return out


def net_byband_full(gpt_flux_dn, gpt_flux_up, band_lims, fast=False):
    """net_byband_full (extensions/mo_fluxes_byband_kernels.F90:56-78)."""
    ncol, nlev, ngpt = np.shape(gpt_flux_dn)
    d, u = _a(gpt_flux_dn, fast), _a(gpt_flux_up, fast); bl = np.ascontiguousarray(band_lims, np.int32)
    out = np.empty((ncol, nlev, len(bl)), _dt(fast))
    lib(fast).orc_net_byband_full(ncol, nlev, ngpt, len(bl), _ip(bl), _p(d), _p(u), _p(out))
    return out


def net_flux(flux_dn, flux_up, fast=False):
    """net_byband_precalc / net_broadband_precalc: down - up."""
    d, u = _a(flux_dn, fast), _a(flux_up, fast)
    out = np.empty_like(d)
    lib(fast).orc_net_flux(C.c_size_t(d.size), _p(d), _p(u), _p(out))
    return out


def compute_optimal_angles(tau, band_lims, optimal_angle_fit, fast=False):
    """compute_optimal_angles (rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758): tau (ncol, nlay, ngpt), optimal_angle_fit
    (nbnd, 2) [== Fortran (2, nbnd)] -> (ncol, ngpt)."""
    ncol, nlay, ngpt = np.shape(tau)
    t, fit = _a(tau, fast), _a(optimal_angle_fit, fast)
    g2b = np.zeros(ngpt, np.int32)
    for b, (s, e) in enumerate(np.asarray(band_lims)):
        g2b[s - 1:e] = b
    out = np.empty((ncol, ngpt), _dt(fast))
    lib(fast).orc_compute_optimal_angles(ncol, nlay, ngpt, _ip(g2b), _p(fit), _p(t), _p(out))
    return out


def set_solar_variability(solar_quiet, solar_facular, solar_sunspot, mg_index, sb_index, tsi=None, fast=False):
    """set_solar_variability (+ set_tsi when tsi is given), rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1120."""
    q, f, s = _a(solar_quiet, fast), _a(solar_facular, fast), _a(solar_sunspot, fast)
    out = np.empty_like(q)
    lib(fast).orc_set_solar_variability(len(q), _p(q), _p(f), _p(s), _sc(mg_index, fast), _sc(sb_index, fast),
                                        int(tsi is not None), _sc(0.0 if tsi is None else tsi, fast), _p(out))
    return out


def solar_var_ind_interp(avgcyc_ind, solcycfrac):
    """ty_solar_var%solar_var_ind_interp, extensions/solar_variability/mo_solar_variability.F90:91-183, restated step by step in
    fp32 (wp = sp).  avgcyc_ind: (nsolarfrac, 2) == the reference's (2, nsolarfrac).  -> (error_msg, mg_index, sb_index)."""
    f = np.float32
    a = np.asarray(avgcyc_ind, f)
    x = f(solcycfrac)
    if x < 0 or x > 1:                                                    # :122-124
        return "solar_var_ind_interp: solcycfrac out of range", None, None
    n = a.shape[0]
    if x == 0:                                                            # :139-141
        return "", a[0, 0], a[0, 1]
    if x == 1:                                                            # :143-145
        return "", a[n - 1, 0], a[n - 1, 1]
    intrvl_len = f(1) / f(n - 2)                                          # :148-149
    hf = f(0.5) * intrvl_len
    if x <= hf:                                                           # :153-157
        sfid, fraclo, frachi = 1, f(0), hf
    if x > hf and x < f(1) - hf:                                          # :160-164
        sfid = int(np.floor(f(x - hf) * f(n - 2))) + 2
        fraclo = f(f(sfid - 2) * intrvl_len) + hf
        frachi = fraclo + intrvl_len
    if x >= f(1) - hf:                                                    # :168-172
        sfid, fraclo, frachi = n - 1, f(1) - hf, f(1)
    intfrac = f(x - fraclo) / f(frachi - fraclo)                          # :175-179 (sfid is 1-based)
    mg = a[sfid - 1, 0] + f(intfrac * f(a[sfid, 0] - a[sfid - 1, 0]))
    sb = a[sfid - 1, 1] + f(intfrac * f(a[sfid, 1] - a[sfid - 1, 1]))
    return "", f(mg), f(sb)


def num_threads(fast=True):
    return int(lib(fast).orc_num_threads())
