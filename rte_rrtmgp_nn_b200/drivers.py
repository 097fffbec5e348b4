"""Replicas of the reference's example drivers on top of the C ABI: the callers and the file formats on
either side of the hot path.

  rrtmgp_rfmip_lw / rrtmgp_rfmip_sw   examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90, rrtmgp_rfmip_sw.F90: 100 sites x 18
                                      experiments in blocks of `block_size` columns, gas optics (NN) + rte, fluxes unblocked
                                      to (expt, site, level) and written as rlu/rld (rsu/rsd) like unblock_and_write
                                      (mo_rfmip_io.F90:734-870) -- netCDF-4, by the library's own writer (rfmip_io, ncio).
  rrtmgp_allsky                       examples/all-sky/rrtmgp_allsky.F90:150-446: Garand atmosphere 1 replicated ncol times,
                                      the cloud recipe of :333-350, LUT or Pade cloud optics, (delta-scaling,) increment, rte;
                                      write_lw_fluxes / write_sw_fluxes (mo_garand_atmos_io.F90:92-170).
The spectral tables are synthetic (the k-distribution files are not in the reference tree), so the numbers are not the
published ones; the data flow, blocking, layouts and file structure are.
"""
import os

import numpy as np

from . import api, rfmip, rfmip_io, spectral
from .ncio import NcFile

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
NN_DIR = os.path.join(ROOT, "data", "nn")
GARAND = os.path.join(ROOT, "data", "garand", "garand-atmos-1.nc")
LW_NETS = ("lw-g256-2018-12-04_absorption_58_58.nc", "lw-g256-2018-12-04_planck_frac_16_16.nc")
SW_NETS = ("sw-g224-2018-12-04-absorption_16_16.nc", "sw-g224-2018-12-04-rayleigh_16_16.nc")


def _ok(error_msg):
    """stop_on_err of the reference drivers (mo_rrtmgp_clr_all_sky / the drivers' own copies): an error message aborts."""
    if error_msg != "":
        raise RuntimeError(error_msg)


def _nets(ctx, files):
    return [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in files]


def write_rfmip_fluxes(path, names, fluxes, nexp, nsite):
    """unblock_and_write (mo_rfmip_io.F90:734-870): (ncol, nlev) column-blocked fluxes -> variables `names` with
    dimensions (expt, site, level), netCDF-4."""
    rfmip_io.unblock_and_write(path, names, fluxes, nexp, nsite)


def write_allsky_fluxes(path, names, fluxes):
    """write_lw_fluxes / write_sw_fluxes (mo_garand_atmos_io.F90:92-170): variables (lev, col), netCDF-4 (a new file; the
    reference adds them to its input file)."""
    ncol, nlev = fluxes[0].shape
    with NcFile(path, "w") as f:
        f.create_dim("col", ncol); f.create_dim("lev", nlev)
        for name, a in zip(names, fluxes):
            f.write_field(name, ("lev", "col"), np.asarray(a, np.float32).T)


def read_atmos(path=GARAND):
    """read_atmos (mo_garand_atmos_io.F90:41-88): the file's variables are (lay|lev, col); returned as (col, lay|lev), this
    package's layout."""
    out = {}
    with NcFile(path) as f:
        rd = lambda k: np.ascontiguousarray(f.read_field(k).T)
        for k in ("p_lay", "t_lay", "p_lev", "t_lev"):
            out[k] = rd(k)
        for g in ("h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "n2"):
            out["vmr_" + g] = rd("vmr_" + g)
        if f.var_exists("col_dry"):
            out["col_dry"] = rd("col_dry")
    return out


def rrtmgp_rfmip_lw(ctx=None, block_size=8, out_path=None, n_quad_angles=1, columns=None):
    """rrtmgp_rfmip_lw.F90:240-446.  Returns flux_up, flux_dn (ncol, nlev) in column order expt*100 + site."""
    ctx = ctx or api.default_context()
    atm = rfmip.load(columns=columns)
    ncol, nlay = atm["play"].shape
    k_dist = api.ty_gas_optics_rrtmgp(ctx); k_dist.load(spectral.synthetic_kdist_lw(256))
    nets = _nets(ctx, LW_NETS)
    up = np.empty((ncol, nlay + 1), np.float32); dn = np.empty_like(up)
    # the block loop of the driver (:368-446): block_size columns per call, here through the host-buffer entry point
    for b0 in range(0, ncol, block_size):
        s = slice(b0, min(b0 + block_size, ncol))
        gc = api.ty_gas_concs()
        for k, v in atm["gases"].items():
            gc.set_vmr(k, v[s])
        u, d = api.lw_fluxes_host(k_dist, nets, atm["play"][s], atm["plev"][s], atm["tlay"][s], atm["tsfc"][s], atm["sfc_emis"][s], gc,
                                  tlev=atm["tlev"][s], n_gauss_angles=n_quad_angles, top_at_1=atm["top_at_1"])
        up[s], dn[s] = u, d
    if out_path and columns is None:
        write_rfmip_fluxes(out_path, ("rlu", "rld"), (up, dn), ncol // 100, 100)
    return up, dn


def rrtmgp_rfmip_sw(ctx=None, block_size=8, out_path=None, columns=None):
    """rrtmgp_rfmip_sw.F90:230-465 (TSI renormalisation, night columns zeroed)."""
    ctx = ctx or api.default_context()
    atm = rfmip.load(columns=columns)
    ncol, nlay = atm["play"].shape
    k_dist = api.ty_gas_optics_rrtmgp(ctx); k_dist.load(spectral.synthetic_kdist_sw(224))
    nets = _nets(ctx, SW_NETS)
    up = np.empty((ncol, nlay + 1), np.float32); dn = np.empty_like(up)
    mu0 = np.where(atm["usecol"], atm["mu0"], -1.0).astype(np.float32)   # night columns: mu0 <= 0 marks them for the library
    for b0 in range(0, ncol, block_size):
        s = slice(b0, min(b0 + block_size, ncol))
        gc = api.ty_gas_concs()
        for k, v in atm["gases"].items():
            gc.set_vmr(k, v[s])
        u, d, _ = api.sw_fluxes_host(k_dist, nets, atm["play"][s], atm["plev"][s], atm["tlay"][s], mu0[s], atm["sfc_alb"][s], gc,
                                     tsi=atm["tsi"][s], top_at_1=atm["top_at_1"])
        up[s], dn[s] = u, d
    if out_path and columns is None:
        write_rfmip_fluxes(out_path, ("rsu", "rsd"), (up, dn), ncol // 100, 100)
    return up, dn


def garand_atmosphere(ncol):
    """read_atmos + `p_lay = spread(p_lay(:,1), ...)` (rrtmgp_allsky.F90:172-195): profile 1 replicated ncol times."""
    z = read_atmos()
    rep = lambda a: np.ascontiguousarray(np.repeat(a[:1], ncol, axis=0))
    atm = dict(play=rep(z["p_lay"]), plev=rep(z["p_lev"]), tlay=rep(z["t_lay"]), tlev=rep(z["t_lev"]))
    atm["gases"] = {g: rep(z["vmr_" + g]) for g in ("h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "n2")}
    atm["top_at_1"] = bool(atm["play"][0, 0] < atm["play"][0, -1])
    return atm


def allsky_clouds(atm, cloud_optics):
    """The cloud recipe of rrtmgp_allsky.F90:333-350."""
    play, tlay = atm["play"], atm["tlay"]
    ncol = play.shape[0]
    icol1 = np.arange(1, ncol + 1)[:, None]
    mask = (play > 100.0 * 100.0) & (play < 900.0 * 100.0) & ((icol1 % 3) != 0)
    t = cloud_optics.tables
    if "radliq_lwr" in t:
        rl = (t["radliq_lwr"], t["radliq_lwr"] + t["liq_step_size"] * (t["liq_nsteps"] - 1))
        ri = (t["radice_lwr"], t["radice_lwr"] + t["ice_step_size"] * (t["ice_nsteps"] - 1))
    else:
        rl = (t["sizreg"][0, 0], t["sizreg"][0, 3]); ri = (t["sizreg"][3, 0], t["sizreg"][3, 3])
    rel_val, rei_val = np.float32(0.5 * (rl[0] + rl[1])), np.float32(0.5 * (ri[0] + ri[1]))
    lwp = np.where(mask & (tlay > 263.0), 10.0, 0.0).astype(np.float32)
    iwp = np.where(mask & (tlay < 273.0), 10.0, 0.0).astype(np.float32)
    return dict(lwp=lwp, iwp=iwp, rel=np.where(lwp > 0, rel_val, 0).astype(np.float32), rei=np.where(iwp > 0, rei_val, 0).astype(np.float32))


def rrtmgp_allsky(ncol, nloops=1, band="lw", ctx=None, use_pade=False, out_path=None):
    """rrtmgp_allsky.F90:150-446 for one band; returns the fluxes of the last loop as numpy arrays (ncol, nlev)."""
    torch = api._torch()
    ctx = ctx or api.default_context()
    atm = garand_atmosphere(ncol)
    nlay = atm["play"].shape[1]
    lw = band == "lw"
    k_dist = api.ty_gas_optics_rrtmgp(ctx)
    k_dist.load(spectral.synthetic_kdist_lw(256) if lw else spectral.synthetic_kdist_sw(224))
    nets = _nets(ctx, LW_NETS if lw else SW_NETS)
    coef = os.path.join(ROOT, "data", "cloud_optics", f"rrtmgp-cloud-optics-coeffs-{band}.nc")
    co = api.ty_cloud_optics(ctx)
    _ok(co.load_pade(**api.load_cloud_pade_file(coef)) if use_pade else co.load(**api.load_cloud_lut_file(coef)))
    cl = allsky_clouds(atm, co)
    gc = api.ty_gas_concs()
    for k, v in atm["gases"].items():
        gc.set_vmr(k, v)
    mk = lambda: torch.zeros((ncol, nlay + 1), device="cuda")
    if lw:
        atmos = api.ty_optical_props_1scl(); _ok(atmos.alloc_1scl(ncol, nlay, k_dist))
        clouds = api.ty_optical_props_1scl(); _ok(clouds.alloc_1scl(ncol, nlay, k_dist, by_band=True))
        src = api.ty_source_func_lw(); _ok(src.alloc(ncol, nlay, k_dist))
        t_sfc = atm["tlev"][:, nlay if atm["top_at_1"] else 0].copy()
        emis = np.full((ncol, k_dist.nband), 0.98, np.float32)
        fl = api.ty_fluxes_broadband(mk(), mk())
    else:
        atmos = api.ty_optical_props_2str(); _ok(atmos.alloc_2str(ncol, nlay, k_dist))
        clouds = api.ty_optical_props_2str(); _ok(clouds.alloc_2str(ncol, nlay, k_dist, by_band=True))
        toa = torch.empty((ncol, k_dist.ngpt), device="cuda")
        alb = np.full((ncol, k_dist.ngpt), 0.06, np.float32)
        mu0 = np.full(ncol, 0.86, np.float32)
        fl = api.ty_fluxes_broadband(mk(), mk(), None, mk())
    for _ in range(nloops):  # :366-446
        _ok(co.cloud_optics(cl["lwp"], cl["iwp"], cl["rel"], cl["rei"], clouds))
        if lw:
            _ok(k_dist.gas_optics(atm["play"], atm["plev"], atm["tlay"], t_sfc, gc, atmos, src, tlev=atm["tlev"], neural_nets=nets))
            _ok(clouds.increment(atmos))
            _ok(api.rte_lw(atmos, atm["top_at_1"], src, emis, fl))
        else:
            _ok(k_dist.gas_optics(atm["play"], atm["plev"], atm["tlay"], gc, atmos, toa, neural_nets=nets))
            _ok(clouds.delta_scale())
            _ok(clouds.increment(atmos))
            _ok(api.rte_sw(atmos, atm["top_at_1"], mu0, toa, alb, alb, fl))
    out = [fl.flux_up.cpu().numpy(), fl.flux_dn.cpu().numpy()] + ([] if lw else [fl.flux_dn_dir.cpu().numpy()])
    if out_path:
        write_allsky_fluxes(out_path, ("lw_flux_up", "lw_flux_dn") if lw else ("sw_flux_up", "sw_flux_dn", "sw_flux_dir"), out)
    return out
