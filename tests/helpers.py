"""Shared set-up for the tests: spectral tables, networks (oracle + device), atmospheres."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NN_DIR = os.path.join(ROOT, "data", "nn")
GOLDEN = os.path.join(ROOT, "tests", "golden")

LW_G256 = ("lw-g256-2018-12-04_absorption_58_58.nc", "lw-g256-2018-12-04_planck_frac_16_16.nc")
SW_G224 = ("sw-g224-2018-12-04-absorption_16_16.nc", "sw-g224-2018-12-04-rayleigh_16_16.nc")
LW_G128 = ("lw-g128-210809_absorption_BEST.nc", "lw-g128-210809_planck_frac_BEST.nc")
LW_G128_BOTH = ("lw-g128-210809_both_BEST.nc",)
LW_G128_NWP = ("rrtmgp-data-lw-g128-210809_NN_GCM_NWP_absorption.nc", "rrtmgp-data-lw-g128-210809_NN_GCM_NWP_planck_frac.nc")
SW_G112 = ("sw-g112-210809_absorption_BEST.nc", "sw-g112-210809_rayleigh_BEST.nc")

# tolerances stated by BASELINE.json north_star
FLUX_TOL = 0.01      # W m-2, every level
HR_TOL = 1.0e-3      # K day-1
TAU_RTOL = 1.0e-4    # relative, fp32 path; applied with an absolute floor (SURVEY.md section 7 "NN precision")


def oracle_nets(files):
    import nc4min
    import oracle as O
    return [O.Net(nc4min.load_nn_model(os.path.join(NN_DIR, f))) for f in files]


def device_nets(ctx, files):
    from rte_rrtmgp_nn_b200 import api
    return [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in files]


def gas_concs(gases):
    from rte_rrtmgp_nn_b200 import api
    gc = api.ty_gas_concs(list(gases.keys()))
    for k, v in gases.items():
        assert gc.set_vmr(k, v) == ""
    return gc


def tau_rel_err(tau, ref):
    """Relative tau error with an absolute floor of 1e-6 x the largest tau of the same sample
    (weak g-points sit at fp32 reorder-noise level: tau = (ystd*z+ymean)^8 amplifies dz by 8*ystd/ymean)."""
    floor = 1e-6 * np.max(np.abs(ref), axis=-1, keepdims=True) + 1e-30
    return np.abs(tau - ref) / np.maximum(np.abs(ref), floor)
