"""RFMIP clear-sky inputs (BASELINE.json configs[0..1]): the 100 sites x 18 experiments = 1800 columns x 60 layers of
examples/rfmip-clear-sky/multiple_input4MIPs_radiation_RFMIP_UColorado-RFMIP-1-2_none.nc, read with rfmip_io (the library's own
netCDF-4 reader) and given the drivers' input conditioning:
  p_lay clipped to press_min, top p_lev := press_min + eps      rrtmgp_rfmip_lw.F90:287, 300-305
  usecol / mu0 (night columns get mu0 = 1 and are zeroed later)  rrtmgp_rfmip_sw.F90:285-287, 428-434
"""
import functools

import numpy as np

from . import rfmip_io

GM_GASES = rfmip_io.GM_GASES


@functools.lru_cache(maxsize=1)
def _file():
    p_lay, p_lev, t_lay, t_lev = rfmip_io.read_and_block_pt()
    sfc_emis, sfc_t = rfmip_io.read_and_block_lw_bc()
    sfc_alb, tsi, sza = rfmip_io.read_and_block_sw_bc()
    gases = rfmip_io.read_and_block_gases_ty()
    return dict(p_lay=p_lay, p_lev=p_lev, t_lay=t_lay, t_lev=t_lev, sfc_emis=sfc_emis, sfc_t=sfc_t, sfc_alb=sfc_alb, tsi=tsi, sza=sza,
                gases=gases)


def load(press_min=1.00518357, columns=None):
    z = _file()
    sel = slice(None) if columns is None else np.asarray(columns)
    f32 = np.float32
    p_lay = z["p_lay"][sel].copy(); p_lev = z["p_lev"][sel].copy()
    pm = f32(press_min)
    p_lay[p_lay < pm] = pm + np.spacing(pm)
    top_at_1 = bool(p_lay[0, 0] < p_lay[0, -1])
    if top_at_1:
        p_lev[:, 0] = pm + np.finfo(f32).eps
    else:
        p_lev[:, -1] = pm + np.finfo(f32).eps
    gases = {g: np.ascontiguousarray(v[sel]).astype(f32) for g, v in z["gases"].items()}
    sza = z["sza"][sel]
    usecol = sza < f32(90.0) - f32(2.0) * np.spacing(f32(90.0))
    mu0 = np.where(usecol, np.cos(sza * f32(np.arccos(-1.0) / 180.0)), f32(1.0)).astype(f32)
    return dict(play=p_lay, plev=p_lev, tlay=z["t_lay"][sel].copy(), tlev=z["t_lev"][sel].copy(), tsfc=z["sfc_t"][sel].copy(),
                sfc_emis=z["sfc_emis"][sel].copy(), sfc_alb=z["sfc_alb"][sel].copy(), mu0=mu0, usecol=usecol,
                tsi=z["tsi"][sel].copy(), gases=gases, top_at_1=top_at_1)
