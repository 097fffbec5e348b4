"""Spectral metadata of the k-distribution that the NN path still needs.

In the reference these tables come from rrtmgp/data/rrtmgp-data-{lw,sw}-g*.nc through
examples/mo_load_coefficients.F90:130-135,184-186,216-221 into ty_gas_optics_rrtmgp%load
(rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326): bnd_limits_gpt, totplnk, temp_ref_min/max,
press_ref_min/max, solar_source.  Those files are absent from the reference snapshot
(.MISSING_LARGE_BLOBS:4-7), so `synthetic_kdist_*` rebuilds physically sensible stand-ins:
the true RRTMGP band wavenumber limits (they are in-tree, in
extensions/cloud_optics/rrtmgp-cloud-optics-coeffs-{lw,sw}.nc: bnd_limits_wavenumber), equal
g-point counts per band, totplnk as band-integrated Planck radiance on a 1 K grid, and a 5778 K
black-body solar spectrum scaled to a TSI.  A real k-distribution file's tables can be passed to
`make_kdist` unchanged.
"""
import numpy as np

LW_BAND_LIMS_WVN = np.array([[10, 250], [250, 500], [500, 630], [630, 700], [700, 820], [820, 980], [980, 1080],
                             [1080, 1180], [1180, 1390], [1390, 1480], [1480, 1800], [1800, 2080], [2080, 2250],
                             [2250, 2390], [2390, 2680], [2680, 3250]], dtype=np.float64)
SW_BAND_LIMS_WVN = np.array([[820, 2680], [2680, 3250], [3250, 4000], [4000, 4650], [4650, 5150], [5150, 6150],
                             [6150, 7700], [7700, 8050], [8050, 12850], [12850, 16000], [16000, 22650],
                             [22650, 29000], [29000, 38000], [38000, 50000]], dtype=np.float64)

_H = 6.62607015e-34
_C = 2.99792458e8
_KB = 1.380649e-23


def _planck_band_radiance(T, lo_cm, hi_cm, n=400):
    """Integral of B_nu(T) d nu over [lo,hi] cm^-1, W m^-2 sr^-1 (float64)."""
    nu = np.linspace(lo_cm, hi_cm, n) * 100.0  # m^-1
    x = _H * _C * nu[None, :] / (_KB * np.asarray(T, np.float64)[:, None])
    b = 2.0 * _H * _C * _C * nu[None, :] ** 3 / np.expm1(x)
    return np.trapezoid(b, nu, axis=1)


def make_kdist(band_lims_wvn, band_lims_gpt, totplnk=None, temp_ref_min=160.0, totplnk_delta=1.0,
               solar_source=None, press_ref_min=1.00518357, press_ref_max=109663.31, temp_ref_max=355.0):
    band_lims_gpt = np.ascontiguousarray(band_lims_gpt, np.int32)
    ngpt = int(band_lims_gpt[-1, 1])
    kd = dict(nbnd=len(band_lims_gpt), ngpt=ngpt, band_lims_wvn=np.asarray(band_lims_wvn, np.float32),
              band_lims_gpt=band_lims_gpt, temp_ref_min=float(temp_ref_min), temp_ref_max=float(temp_ref_max),
              totplnk_delta=float(totplnk_delta), press_ref_min=float(press_ref_min),
              press_ref_max=float(press_ref_max))
    gpt2band = np.zeros(ngpt, np.int32)
    for b, (s, e) in enumerate(band_lims_gpt):
        gpt2band[s - 1:e] = b + 1
    kd["gpt2band"] = gpt2band
    if totplnk is not None:
        kd["totplnk"] = np.ascontiguousarray(totplnk, np.float32)  # [nbnd][nPlanckTemp]
    if solar_source is not None:
        kd["solar_source"] = np.ascontiguousarray(solar_source, np.float32)
    return kd


def _even_gpt_limits(nbnd, ngpt):
    per = ngpt // nbnd
    assert per * nbnd == ngpt
    return np.array([[b * per + 1, (b + 1) * per] for b in range(nbnd)], np.int32)


def synthetic_kdist_lw(ngpt=256, ntemp=196):
    lims = LW_BAND_LIMS_WVN
    T = 160.0 + np.arange(ntemp, dtype=np.float64)
    tot = np.stack([_planck_band_radiance(T, lo, hi) for lo, hi in lims])  # [nbnd][ntemp]
    return make_kdist(lims, _even_gpt_limits(len(lims), ngpt), totplnk=tot, temp_ref_min=160.0, totplnk_delta=1.0)


def synthetic_kdist_sw(ngpt=224, tsi=1361.0):
    lims = SW_BAND_LIMS_WVN
    gl = _even_gpt_limits(len(lims), ngpt)
    band = np.array([_planck_band_radiance(np.array([5778.0]), lo, hi, 2000)[0] for lo, hi in lims])
    per = ngpt // len(lims)
    # descending weights inside a band so g-points are not all identical
    w = np.linspace(1.6, 0.4, per); w /= w.sum()
    src = (band[:, None] * w[None, :]).ravel()
    kd = make_kdist(lims, gl, solar_source=src)
    set_tsi(kd, tsi)
    return kd


def set_tsi(kd, tsi):
    """ty_gas_optics_rrtmgp%set_tsi: rrtmgp/mo_gas_optics_rrtmgp.F90:1097-1120 (fp32 arithmetic)."""
    s = kd["solar_source"].astype(np.float32)
    norm = np.float32(1.0) / np.sum(s, dtype=np.float32)
    kd["solar_source"] = (s * np.float32(tsi) * norm).astype(np.float32)
    return kd
