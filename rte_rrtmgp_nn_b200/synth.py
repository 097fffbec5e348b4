"""Synthetic atmospheres for the GCM-scale configs (BASELINE.json configs[2..4]).

Profiles are drawn inside the NN training ranges (input min/max stored in the weight files):
T 160-320.5 K, ln p 5.15e-3..11.55, h2o^(1/4) 0.0101..0.508, o3^(1/4) 4.36e-3..0.0632.
top_at_1 = True (index 0 is the model top) unless flipped by the caller.
"""
import numpy as np

# present-day minor-gas mixing ratios, rrtmgp/mo_gas_ref_concentrations.F90:44-58 magnitudes
MINOR_GASES = dict(cfc11=2.3e-10, cfc12=5.2e-10, co=1.2e-7, ccl4=8.0e-11, cfc22=2.3e-10, hfc143a=1.5e-11,
                   hfc125=1.5e-11, hfc23=2.7e-11, hfc32=8.0e-12, hfc134a=8.0e-11, cf4=8.0e-11)


def make_atmosphere(ncol, nlay, seed=12345, dtype=np.float32):
    rng = np.random.default_rng(seed)
    psfc = rng.uniform(95000.0, 105000.0, size=(ncol, 1))
    # hybrid-like level grid from 1 Pa (top, index 0) to p_sfc
    eta = np.linspace(0.0, 1.0, nlay + 1) ** 2.2
    plev = 1.0 + eta[None, :] * (psfc - 1.0)
    play = 0.5 * (plev[:, 1:] + plev[:, :-1])
    # temperature: troposphere lapse + isothermal stratosphere + warm upper layers, plus noise
    z = -7000.0 * np.log(np.maximum(play, 1.0) / psfc)
    tsurf = rng.uniform(250.0, 305.0, size=(ncol, 1))
    t = np.maximum(tsurf - 6.5e-3 * z, 212.0)
    t = t + np.clip((z - 20000.0) * 1.6e-3, 0.0, 55.0)
    t = np.where(z > 50000.0, np.maximum(t - (z - 50000.0) * 2.2e-3, 180.0), t)
    t = np.clip(t + rng.normal(0.0, 3.0, size=t.shape), 165.0, 315.0)
    # level temperatures by simple averaging + surface extrapolation (kept inside the range)
    tlev = np.empty((ncol, nlay + 1))
    tlev[:, 1:-1] = 0.5 * (t[:, 1:] + t[:, :-1])
    tlev[:, 0] = t[:, 0]
    tlev[:, -1] = np.clip(t[:, -1] + 1.0, 165.0, 315.0)
    tsfc = np.clip(tlev[:, -1] + rng.normal(0.0, 2.0, size=ncol), 165.0, 318.0)
    # water vapour decreasing with height, ozone peaking aloft
    q0 = 10.0 ** rng.uniform(-3.0, -1.6, size=(ncol, 1))
    h2o = np.clip(q0 * (play / psfc) ** 3.0, 1.2e-6, 3.0e-2)
    o3 = 1.0e-8 + 8.0e-6 * np.exp(-0.5 * ((np.log(play) - np.log(1000.0)) / 1.1) ** 2)
    o3 = np.clip(o3 * rng.uniform(0.7, 1.3, size=(ncol, 1)), 2.0e-9, 1.2e-5)
    gases = dict(h2o=h2o.astype(dtype), o3=o3.astype(dtype),
                 co2=np.float32(rng.uniform(2.8e-4, 1.1e-3)), ch4=np.float32(rng.uniform(8e-7, 2.5e-6)),
                 n2o=np.float32(rng.uniform(2.7e-7, 3.9e-7)))
    for k, v in MINOR_GASES.items():
        gases[k] = np.float32(v)
    return dict(play=play.astype(dtype), plev=plev.astype(dtype), tlay=t.astype(dtype), tlev=tlev.astype(dtype),
                tsfc=tsfc.astype(dtype), gases=gases,
                sfc_emis=np.full(ncol, 0.98, dtype), sfc_alb=rng.uniform(0.05, 0.6, size=ncol).astype(dtype),
                mu0=rng.uniform(0.05, 1.0, size=ncol).astype(dtype), top_at_1=True)


def flip_vertical(atm):
    """Same atmosphere ordered bottom-to-top (top_at_1 = False)."""
    out = dict(atm)
    for k in ("play", "plev", "tlay", "tlev"):
        out[k] = np.ascontiguousarray(atm[k][:, ::-1])
    g = dict(atm["gases"])
    for k, v in g.items():
        if np.ndim(v) == 2:
            g[k] = np.ascontiguousarray(v[:, ::-1])
    out["gases"] = g
    out["top_at_1"] = not atm["top_at_1"]
    return out


def make_clouds(atm, seed=7):
    """The all-sky example's recipe: examples/all-sky/rrtmgp_allsky.F90:333-350 -- cloud where
    100 hPa < p < 900 hPa and mod(icol,3) /= 0 (1-based icol); LWP = IWP = 10 g/m2, re mid-range."""
    play = atm["play"]
    ncol, nlay = play.shape
    icol1 = np.arange(1, ncol + 1)[:, None]
    mask = (play > 100.0 * 100.0) & (play < 900.0 * 100.0) & ((icol1 % 3) != 0)
    rel_val = np.float32(0.5 * (2.5 + 21.5)); rei_val = np.float32(0.5 * (10.0 + 180.0))
    lwp = np.where(mask, 10.0, 0.0).astype(np.float32)
    iwp = np.where(mask, 10.0, 0.0).astype(np.float32)
    rel = np.where(mask, rel_val, 0.0).astype(np.float32)
    rei = np.where(mask, rei_val, 0.0).astype(np.float32)
    return dict(lwp=lwp, iwp=iwp, rel=rel, rei=rei)
