"""Golden vectors computed BY THE REFERENCE'S OWN PYTHON (run in the build container only; /root/reference does not exist on
the GPU box).  tests/golden/ref_python_golden.npz is what pins the parts of the path that the reference also holds in Python:

  get_col_dry                      examples/rrtmgp-nn-training/ml_load_save_preproc.py:283-293
                                   (the twin of rrtmgp/mo_gas_optics_rrtmgp.F90 get_col_dry)
  preproc_minmax_inputs_rrtmgp     ml_load_save_preproc.py:416-435 (log p, h2o^(1/4), o3^(1/4), min-max scaling: the twin of
                                   compute_nn_inputs, rrtmgp/mo_gas_optics_rrtmgp.F90:708-760)
  preproc_pow_standardization_reverse  ml_load_save_preproc.py:329-340 ((ystd z + ymean)^8; times col_dry it is the post-processing
                                   of output_sgemm_tau, neural/mod_network_rrtmgp.F90:125-236)
  calc_heatingrates                ml_eval_funcs.py:23-34 (K/day; twin of rrtmgp_lw_eval_nn_rfmip.F90:623-651)
  ymeans_* / ysigma_*              ml_scaling_coefficients.py:30-190 (the constants the shipped 2018 weight files carry as
                                   nn_output_coeffs_mean / _std)

The reference's modules are imported UNMODIFIED from /root/reference; the two imports they need and this image lacks (netCDF4, only
used by their file readers / writers, and matplotlib, only used by their plots) are stubbed with empty modules.  numba is present,
so the @njit functions run as written.  The network's raw outputs z, which the reference computes with Keras (absent here), come
from a float64 numpy evaluation of the shipped weights (the same three mat-muls + softsign; written out below): what is pinned is
everything AROUND the mat-muls.  The RTE solvers exist only in Fortran in the reference: they stay unpinned.

Inputs: the reference's RFMIP profiles (tests/golden/rfmip_inputs.npz, tools/make_rfmip_fixture.py), every 50th column.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/examples/rrtmgp-nn-training"
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def import_reference_python():
    for name in ("netCDF4", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    for name, attr in (("netCDF4", "Dataset"), ("matplotlib.colors", "LogNorm"), ("mpl_toolkits.axes_grid1", "make_axes_locatable")):
        if not hasattr(sys.modules[name], attr):
            setattr(sys.modules[name], attr, None)
    sys.path.insert(0, REF)
    import ml_eval_funcs, ml_load_save_preproc, ml_scaling_coefficients
    return ml_load_save_preproc, ml_scaling_coefficients, ml_eval_funcs


MODELS = {  # tag -> (file under data/nn, the reference's constants for it)
    "sw_abs": ("sw-g224-2018-12-04-absorption_16_16.nc", "ymeans_sw_absorption_224", "ysigma_sw_absorption_224"),
    "sw_ray": ("sw-g224-2018-12-04-rayleigh_16_16.nc", "ymeans_sw_ray_224", "ysigma_sw_ray_224"),
    "lw_abs": ("lw-g256-2018-12-04_absorption_58_58.nc", "ymeans_lw_absorption_256", "ysigma_lw_absorption_256"),
    # the 2021 generation: its INPUT scaling is the reference's xmin_all / xmax_all (ml_scaling_coefficients.py:14-28), which the 16 shipped
    # lw-g128 files carry to the bit; its output scaling exists in the weight file only (None: taken from the file)
    "lw128_abs": ("lw-g128-210809_absorption_BEST.nc", None, None),
}


def main():
    P, SC, E = import_reference_python()
    from nc4min import load_nn_model
    d = np.load(os.path.join(ROOT, "tests", "golden", "rfmip_inputs.npz"))
    cols = np.arange(0, 1800, 50)   # 36 columns: two sites of each of the 18 experiments
    out = {"columns": cols}
    h2o, o3, plev, play, tlay = (d[k][cols] for k in ("h2o", "o3", "p_lev", "p_lay", "t_lay"))
    ncol, nlay = play.shape
    col_dry = P.get_col_dry(h2o, plev)
    out["col_dry"] = col_dry
    for tag, (fn, ym_name, ys_name) in MODELS.items():
        m = load_nn_model(os.path.join(ROOT, "data", "nn", fn))
        names = m["input_names"]
        if ym_name is None:
            ymean, ysigma = m["ymean"].astype(np.float64), m["ystd"].astype(np.float64)
            xco = tuple(np.asarray(c, np.float32)[[SC.input_names_all.index(n) for n in names]] for c in SC.xcoeffs_all)
            out[tag + "_xmin"], out[tag + "_xmax"] = xco
        else:
            ymean, ysigma = getattr(SC, ym_name).astype(np.float64), getattr(SC, ys_name).astype(np.float64)
            out[tag + "_ymean"] = np.asarray(getattr(SC, ym_name))
            out[tag + "_ysigma"] = np.asarray(getattr(SC, ys_name))
            xco = (m["xmin"], m["xmax"])
        # the raw inputs, one row per (column, layer), in the order of the model's own input names
        x_raw = np.empty((ncol * nlay, len(names)), np.float32)
        for i, n in enumerate(names):
            if n == "tlay": v = tlay
            elif n == "play": v = play
            elif n == "h2o": v = h2o
            elif n == "o3": v = o3
            else: v = np.broadcast_to(d["gm_" + n][cols][:, None], (ncol, nlay))
            x_raw[:, i] = np.asarray(v, np.float32).reshape(-1)
        x = P.preproc_minmax_inputs_rrtmgp(x_raw, xco)
        out[tag + "_nn_inputs"] = x.reshape(ncol, nlay, -1)
        a = x.astype(np.float64)
        for l in range(3):
            a = a @ m["W"][l].astype(np.float64) + m["b"][l].astype(np.float64)
            if l < 2:
                a = a / (np.abs(a) + 1)
        y = P.preproc_pow_standardization_reverse(a, 8, ymean, ysigma)
        out[tag + "_tau"] = (y * col_dry.reshape(-1, 1).astype(np.float64)).reshape(ncol, nlay, -1).astype(np.float32)   # (stored rounded to fp32)
    # heating rates of a smooth synthetic flux profile on the RFMIP pressure levels
    rng = np.random.default_rng(7)
    fdn = np.cumsum(rng.uniform(0.5, 8.0, size=plev.shape), 1)
    fup = 420.0 - np.cumsum(rng.uniform(0.2, 5.0, size=plev.shape), 1)[:, ::-1]
    out["hr_flux_up"], out["hr_flux_dn"] = fup.astype(np.float32), fdn.astype(np.float32)
    out["hr_K_day"] = E.calc_heatingrates(out["hr_flux_up"].astype(np.float64), out["hr_flux_dn"].astype(np.float64), plev.astype(np.float64))[0]
    dst = os.path.join(ROOT, "tests", "golden", "ref_python_golden.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst) / 1e6, "MB", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
