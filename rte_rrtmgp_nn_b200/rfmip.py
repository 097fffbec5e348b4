"""RFMIP clear-sky inputs (BASELINE.json configs[0..1]): the 100 sites x 18 experiments = 1800 columns x 60 layers
of examples/rfmip-clear-sky/multiple_input4MIPs_radiation_RFMIP_UColorado-RFMIP-1-2_none.nc, extracted once into
tests/golden/rfmip_inputs.npz by tools/make_rfmip_fixture.py, with the drivers' input conditioning:
  p_lay clipped to press_min, top p_lev := press_min + eps      rrtmgp_rfmip_lw.F90:287, 300-305
  usecol / mu0 (night columns get mu0 = 1 and are zeroed later)  rrtmgp_rfmip_sw.F90:285-287, 428-434
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "rfmip_inputs.npz")
GM_GASES = ("co2", "n2o", "ch4", "co", "ccl4", "cfc22", "cfc11", "cfc12", "hfc143a", "hfc125", "hfc23", "hfc32", "hfc134a", "cf4")


def load(press_min=1.00518357, columns=None):
    z = np.load(FIXTURE)
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
    ncol, nlay = p_lay.shape
    gases = dict(h2o=z["h2o"][sel].copy(), o3=z["o3"][sel].copy())
    for g in GM_GASES:
        gases[g] = np.ascontiguousarray(np.broadcast_to(z["gm_" + g][sel][:, None], (ncol, nlay))).astype(f32)
    sza = z["sza"][sel]
    usecol = sza < f32(90.0) - f32(2.0) * np.spacing(f32(90.0))
    mu0 = np.where(usecol, np.cos(sza * f32(np.arccos(-1.0) / 180.0)), f32(1.0)).astype(f32)
    return dict(play=p_lay, plev=p_lev, tlay=z["t_lay"][sel].copy(), tlev=z["t_lev"][sel].copy(), tsfc=z["sfc_t"][sel].copy(),
                sfc_emis=z["sfc_emis"][sel].copy(), sfc_alb=z["sfc_alb"][sel].copy(), mu0=mu0, usecol=usecol,
                tsi=z["tsi"][sel].copy(), gases=gases, top_at_1=top_at_1)
