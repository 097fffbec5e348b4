"""ctypes front end of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  "PARITY UNPINNED" (see oracle.c header): the reference Fortran cannot be built in this image.

The orchestration functions at the bottom restate the reference's callers:
  gas_optics_lw  <- rrtmgp/mo_gas_optics_rrtmgp.F90:239-428 (NN branch :368-411)
  gas_optics_sw  <- rrtmgp/mo_gas_optics_rrtmgp.F90:433-602 (NN branch :529-573, toa :594-599)
  rte_lw         <- rte/mo_rte_lw.F90:60-424
  rte_sw         <- rte/mo_rte_sw.F90:48-266
All arrays are numpy float32, C-order, g-point fastest: tau[ncol, nlay, ngpt].
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


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_libs = {}


def lib(fast=False):
    key = "fast" if fast else "strict"
    if key not in _libs:
        build()
        name = "liboracle_fast.so" if fast else "liboracle.so"
        _libs[key] = C.CDLL(os.path.join(_HERE, "_build", name))
    return _libs[key]


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


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

    def _args(self):
        return (C.c_int(self.nlayers), self.dims_a.ctypes.data_as(C.POINTER(C.c_int)), _fp(self.wpack),
                _fp(self.bpack), self.act.ctypes.data_as(C.POINTER(C.c_int)))


def get_col_dry(vmr_h2o, plev, fast=False):
    ncol, nlay = vmr_h2o.shape
    h, hp = _f(vmr_h2o); p, pp = _f(plev)
    out = np.empty((ncol, nlay), np.float32)
    lib(fast).orc_get_col_dry(ncol, nlay, hp, pp, _fp(out))
    return out


def interp_tlev(play, plev, tlay, fast=False):
    ncol, nlay = play.shape
    a, ap = _f(play); b, bp = _f(plev); c, cp = _f(tlay)
    out = np.empty((ncol, nlay + 1), np.float32)
    lib(fast).orc_interp_tlev(ncol, nlay, ap, bp, cp, _fp(out))
    return out


def compute_nn_inputs(net, play, tlay, gases, fast=False):
    """gases: dict name -> scalar | (nlay,) | (ncol,nlay) array (ty_gas_concs semantics,
    rrtmgp/mo_gas_concentrations.F90:50-88).  Matching is by name (compute_nn_inputs :708-760)."""
    ncol, nlay = play.shape
    nx = net.dims[0]
    ptrs = (C.POINTER(C.c_float) * nx)()
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
            v = np.ascontiguousarray(np.broadcast_to(v, (ncol, nlay)))
            modes[i] = 2
        elif v.ndim == 0 or v.size == 1:
            v = v.reshape(1); modes[i] = 0
        elif v.ndim == 1:
            assert v.shape == (nlay,); modes[i] = 1
        else:
            assert v.shape == (ncol, nlay); modes[i] = 2
        v = np.ascontiguousarray(v)
        keep.append(v)
        ptrs[i] = _fp(v)
    p, pp = _f(play); t, tp = _f(tlay)
    out = np.empty((ncol, nlay, nx), np.float32)
    lib(fast).orc_compute_nn_inputs(ncol, nlay, nx, pp, tp, ptrs, modes.ctypes.data_as(C.POINTER(C.c_int)),
                                    _fp(net.xmin), _fp(net.xmax), _fp(out))
    return out


def output_sgemm_tau(net, x, coldry, output2=None, fast=False):
    x, xp = _f(x.reshape(-1, net.dims[0]))
    nb = x.shape[0]
    cd, cdp = _f(coldry.reshape(-1))
    out = np.empty((nb, net.dims[-1]), np.float32)
    o2p = None
    if output2 is not None:
        assert output2.dtype == np.float32 and output2.flags.c_contiguous
        o2p = _fp(output2)
    lib(fast).orc_output_sgemm_tau(*net._args(), _fp(net.ymean), _fp(net.ystd), nb, xp, cdp, _fp(out), o2p)
    return out


def output_sgemm_pfrac(net, x, fast=False):
    x, xp = _f(x.reshape(-1, net.dims[0]))
    nb = x.shape[0]
    out = np.empty((nb, net.dims[-1]), np.float32)
    lib(fast).orc_output_sgemm_pfrac(*net._args(), nb, xp, _fp(out))
    return out


def output_sgemm_lw(net, x, fast=False):
    x, xp = _f(x.reshape(-1, net.dims[0]))
    nb = x.shape[0]
    out = np.empty((nb, net.dims[-1]), np.float32)
    lib(fast).orc_output_sgemm_lw(*net._args(), nb, xp, _fp(out))
    return out


def predict_nn_lw(nets, nn_inputs, col_dry, ngpt, fast=False):
    """predict_nn_lw_blas_sp: rrtmgp/kernels/mo_gas_optics_kernels.F90:690-774 -> (tau, pfrac)."""
    ncol, nlay, nx = nn_inputs.shape
    if len(nets) == 2:
        tau = output_sgemm_tau(nets[0], nn_inputs, col_dry, fast=fast)
        pfrac = output_sgemm_pfrac(nets[1], nn_inputs, fast=fast)
    else:
        both = output_sgemm_lw(nets[0], nn_inputs, fast=fast)
        tau = np.empty((ncol * nlay, ngpt), np.float32)
        pfrac = np.empty((ncol * nlay, ngpt), np.float32)
        cd, cdp = _f(col_dry.reshape(-1))
        lib(fast).orc_split_both(ncol, nlay, ngpt, _fp(both), _fp(np.ascontiguousarray(nets[0].ymean[:ngpt])),
                                 _fp(np.ascontiguousarray(nets[0].ystd[:ngpt])), cdp, _fp(tau), _fp(pfrac))
    return tau.reshape(ncol, nlay, ngpt), pfrac.reshape(ncol, nlay, ngpt)


def predict_nn_sw(nets, nn_inputs, col_dry, fast=False):
    """predict_nn_sw_blas_sp: rrtmgp/kernels/mo_gas_optics_kernels.F90:869-953 -> (tau_tot, ssa)."""
    ncol, nlay, nx = nn_inputs.shape
    tau = output_sgemm_tau(nets[0], nn_inputs, col_dry, fast=fast)           # tau_abs
    ssa = output_sgemm_tau(nets[1], nn_inputs, col_dry, output2=tau, fast=fast)  # tau -> tau_tot, ssa
    ngpt = tau.shape[-1]
    return tau.reshape(ncol, nlay, ngpt), ssa.reshape(ncol, nlay, ngpt)


def planck_source_nn(kd, tlay, tlev, tsfc, sfc_lay, pfrac, fast=False):
    """compute_Planck_source_nn; returns (sfc_source, sfc_source_Jac, lay_source, lev_source)."""
    ncol, nlay, ngpt = pfrac.shape
    lay = np.array(pfrac, dtype=np.float32, order="C", copy=True)
    lev = np.empty((ncol, nlay + 1, ngpt), np.float32)
    sfc = np.empty((ncol, ngpt), np.float32)
    jac = np.empty((ncol, ngpt), np.float32)
    a, ap = _f(tlay); b, bp = _f(tlev); c, cp = _f(tsfc)
    bl, blp = _i(kd["band_lims_gpt"])
    tp, tpp = _f(kd["totplnk"])
    nbnd, ntemp = tp.shape
    lib(fast).orc_planck_source_nn(ncol, nlay, nbnd, ngpt, ntemp, ap, bp, cp, int(sfc_lay), blp,
                                   C.c_float(kd["temp_ref_min"]), C.c_float(kd["totplnk_delta"]), tpp, _fp(sfc),
                                   _fp(jac), _fp(lay), _fp(lev))
    return sfc, jac, lay, lev


def gas_optics_lw(kd, nets, play, plev, tlay, tsfc, gases, tlev=None, fast=False):
    """ty_gas_optics_rrtmgp%gas_optics (LW, neural_nets present)."""
    ncol, nlay = play.shape
    if tlev is None:
        tlev = interp_tlev(play, plev, tlay, fast)
    h2o = np.ascontiguousarray(np.broadcast_to(np.asarray(gases["h2o"], np.float32), (ncol, nlay)))
    col_dry = get_col_dry(h2o, plev, fast)
    x = compute_nn_inputs(nets[0], play, tlay, gases, fast)
    tau, pfrac = predict_nn_lw(nets, x, col_dry, kd["ngpt"], fast)
    sfc_lay = 1 if play[0, 0] > play[0, nlay - 1] else nlay
    sfc, jac, lay, lev = planck_source_nn(kd, tlay, tlev, tsfc, sfc_lay, pfrac, fast)
    return dict(tau=tau, lay_source=lay, lev_source=lev, sfc_source=sfc, sfc_source_Jac=jac, col_dry=col_dry,
                nn_inputs=x, tlev=tlev)


def gas_optics_sw(kd, nets, play, plev, tlay, gases, fast=False):
    """ty_gas_optics_rrtmgp%gas_optics (SW, neural_nets present, 2str)."""
    ncol, nlay = play.shape
    h2o = np.ascontiguousarray(np.broadcast_to(np.asarray(gases["h2o"], np.float32), (ncol, nlay)))
    col_dry = get_col_dry(h2o, plev, fast)
    x = compute_nn_inputs(nets[0], play, tlay, gases, fast)
    tau, ssa = predict_nn_sw(nets, x, col_dry, fast)
    g = np.zeros_like(tau)
    toa = np.ascontiguousarray(np.broadcast_to(np.asarray(kd["solar_source"], np.float32), (ncol, kd["ngpt"])))
    return dict(tau=tau, ssa=ssa, g=g, toa_src=toa, col_dry=col_dry, nn_inputs=x)


def expand(band_lims_gpt, ngpt, arr):
    ncol, nband = arr.shape
    a, ap = _f(arr); bl, blp = _i(band_lims_gpt)
    out = np.empty((ncol, ngpt), np.float32)
    lib().orc_expand(nband, ngpt, ncol, blp, ap, _fp(out))
    return out


def lw_solver_noscat_GaussQuad(top_at_1, nmus, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source,
                               inc_flux=None, fast=False):
    ncol, nlay, ngpt = tau.shape
    if inc_flux is None:
        inc_flux = np.zeros((ncol, ngpt), np.float32)
    Ds = np.ascontiguousarray(GAUSS_DS[nmus - 1, :nmus]); wts = np.ascontiguousarray(GAUSS_WTS[nmus - 1, :nmus])
    a = [_f(v) for v in (inc_flux, tau, lay_source, lev_source, sfc_emis_gpt, sfc_source)]
    up = np.empty((ncol, nlay + 1), np.float32); dn = np.empty((ncol, nlay + 1), np.float32)
    lib(fast).orc_lw_solver_noscat_GaussQuad(ngpt, nlay, ncol, int(bool(top_at_1)), nmus, _fp(Ds), _fp(wts),
                                             a[0][1], a[1][1], a[2][1], a[3][1], a[4][1], a[5][1], _fp(up), _fp(dn))
    return up, dn


def rte_lw(kd, top_at_1, tau, lay_source, lev_source, sfc_source, sfc_emis, n_gauss_angles=1, inc_flux=None,
           fast=False):
    """rte_lw for ty_optical_props_1scl: expand emissivity by band, zero incident flux, GaussQuad solver."""
    ngpt = tau.shape[-1]
    emis_gpt = expand(kd["band_lims_gpt"], ngpt, sfc_emis)
    return lw_solver_noscat_GaussQuad(top_at_1, n_gauss_angles, tau, lay_source, lev_source, emis_gpt, sfc_source,
                                      inc_flux, fast)


def sw_solver_2stream(top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif, fast=False):
    ncol, nlay, ngpt = tau.shape
    a = [_f(v) for v in (inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir, alb_dif)]
    up = np.empty((ncol, nlay + 1), np.float32); dn = np.empty_like(up); dr = np.empty_like(up)
    lib(fast).orc_sw_solver_2stream(ngpt, nlay, ncol, int(bool(top_at_1)), *[v[1] for v in a], _fp(up), _fp(dn),
                                    _fp(dr))
    return up, dn, dr


def rte_sw(top_at_1, mu0, inc_flux, alb_dir_gpt, alb_dif_gpt, tau, ssa, g, inc_flux_dif=None, fast=False):
    if inc_flux_dif is None:
        inc_flux_dif = np.zeros_like(inc_flux)
    return sw_solver_2stream(top_at_1, inc_flux, inc_flux_dif, tau, ssa, g, mu0, alb_dir_gpt, alb_dif_gpt, fast)


def cloud_optics_lut(co, clwp, ciwp, reliq, reice, two_stream, fast=False):
    """co: dict from rte_rrtmgp_nn_b200.cloud_optics.load_cloud_lut (tables [nbnd][nsteps], ice roughness chosen)."""
    ncol, nlay = clwp.shape
    nbnd = co["extliq"].shape[0]
    a = [_f(v) for v in (clwp, ciwp, reliq, reice)]
    t = [_f(co[k]) for k in ("extliq", "ssaliq", "asyliq", "extice", "ssaice", "asyice")]
    tau = np.empty((ncol, nlay, nbnd), np.float32); ssa = np.empty_like(tau); g = np.empty_like(tau)
    lib(fast).orc_cloud_optics_lut(ncol, nlay, nbnd, a[0][1], a[1][1], a[2][1], a[3][1],
                                   int(co["liq_nsteps"]), C.c_float(co["liq_step_size"]), C.c_float(co["radliq_lwr"]),
                                   t[0][1], t[1][1], t[2][1],
                                   int(co["ice_nsteps"]), C.c_float(co["ice_step_size"]), C.c_float(co["radice_lwr"]),
                                   t[3][1], t[4][1], t[5][1], int(two_stream), _fp(tau), _fp(ssa), _fp(g))
    if two_stream:
        return tau, ssa, g
    return tau


def delta_scale_2str(tau, ssa, g):
    tau = np.array(tau, np.float32, copy=True); ssa = np.array(ssa, np.float32, copy=True); g = np.array(g, np.float32, copy=True)
    lib().orc_delta_scale_2str(C.c_size_t(tau.size), _fp(tau), _fp(ssa), _fp(g))
    return tau, ssa, g


def inc_1scalar_by_1scalar_bybnd(tau1, tau2, gpt_lims):
    ncol, nlay, ngpt = tau1.shape
    tau1 = np.array(tau1, np.float32, copy=True)
    t2, t2p = _f(tau2); gl, glp = _i(gpt_lims)
    lib().orc_inc_1scalar_by_1scalar_bybnd(ngpt, nlay, ncol, _fp(tau1), t2p, t2.shape[-1], glp)
    return tau1


def inc_2stream_by_2stream_bybnd(tau1, ssa1, g1, tau2, ssa2, g2, gpt_lims):
    ncol, nlay, ngpt = tau1.shape
    tau1 = np.array(tau1, np.float32, copy=True); ssa1 = np.array(ssa1, np.float32, copy=True); g1 = np.array(g1, np.float32, copy=True)
    a = [_f(v) for v in (tau2, ssa2, g2)]; gl, glp = _i(gpt_lims)
    lib().orc_inc_2stream_by_2stream_bybnd(ngpt, nlay, ncol, _fp(tau1), _fp(ssa1), _fp(g1), a[0][1], a[1][1], a[2][1],
                                           a[0][0].shape[-1], glp)
    return tau1, ssa1, g1


def heating_rate(flux_up, flux_dn, plev):
    ncol, nlev = flux_up.shape
    a = [_f(v) for v in (flux_up, flux_dn, plev)]
    out = np.empty((ncol, nlev - 1), np.float32)
    lib().orc_heating_rate(ncol, nlev - 1, a[0][1], a[1][1], a[2][1], _fp(out))
    return out


def calc_heating_rate(flux_up, flux_dn, plev):
    ncol, nlev = flux_up.shape
    a = [_f(v) for v in (flux_up, flux_dn, plev)]
    out = np.empty((ncol, nlev - 1), np.float32)
    lib().orc_calc_heating_rate(ncol, nlev - 1, a[0][1], a[1][1], a[2][1], _fp(out))
    return out


def num_threads(fast=True):
    return int(lib(fast).orc_num_threads())
