"""A/B check of the tensor-core gas optics against the fp32 FFMA kernel on the same device inputs (run under gpurun).

    python tools/tc_check.py [ncol] [nlay]

Prints, per output array, the largest relative difference (floored at 1e-4 of the per-sample maximum) and where it is,
for ragged shapes that exercise tiles straddling column boundaries and the clipped last tile."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import helpers as H
from rte_rrtmgp_nn_b200 import api, spectral, synth

ctx = api.default_context(0)


def rel(a, b):
    fl = 1e-4 * np.max(np.abs(b), axis=-1, keepdims=True) + 1e-300
    return np.abs(a - b) / np.maximum(np.abs(b), fl)


def report(name, a, b):
    bad = ~np.isfinite(a)
    e = rel(np.where(bad, 0, a), b)
    i = np.unravel_index(np.argmax(e), e.shape)
    print(f"  {name:12s} shape {a.shape} max rel diff {e.max():.3e} at {i} (tc {a[i]:.6e} ffma {b[i]:.6e}) nonfinite {int(bad.sum())}"
          f" rows>1e-3: {int((e.max(axis=-1) > 1e-3).sum())}")
    return e.max()


def run_lw(ncol, nlay, files, ngpt, flip=False, scalar_h2o=False):
    atm = synth.make_atmosphere(ncol, nlay, seed=5)
    if flip:
        for k in ("play", "plev", "tlay", "tlev"):
            atm[k] = np.ascontiguousarray(atm[k][:, ::-1])
        atm["gases"] = {k: (np.ascontiguousarray(v[:, ::-1]) if np.ndim(v) == 2 else v) for k, v in atm["gases"].items()}
    kd = spectral.synthetic_kdist_lw(ngpt)
    k_lw = api.ty_gas_optics_rrtmgp(ctx); k_lw.load(kd)
    nets = H.device_nets(ctx, files)
    out = {}
    for tcf in (0, 1):
        ctx.set_flag("nn_tensor_cores", tcf)
        op = api.ty_optical_props_1scl(); op.alloc_1scl(ncol, nlay, k_lw)
        src = api.ty_source_func_lw(); src.alloc(ncol, nlay, k_lw)
        for t in (op.tau, src.lay_source, src.lev_source, src.sfc_source, src.sfc_source_Jac):
            t.fill_(-777.0)
        msg = k_lw.gas_optics(atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], H.gas_concs(atm["gases"]), op, src, tlev=atm["tlev"], neural_nets=nets)
        assert msg == "", msg
        torch.cuda.synchronize()
        out[tcf] = {k: getattr(o, k).cpu().numpy() for o, k in ((op, "tau"), (src, "lay_source"), (src, "lev_source"), (src, "sfc_source"), (src, "sfc_source_Jac"))}
    print(f"LW ncol {ncol} nlay {nlay} ngpt {ngpt} flip {flip}:")
    return max(report(k, out[1][k], out[0][k]) for k in out[0])


def run_sw(ncol, nlay, files, ngpt):
    atm = synth.make_atmosphere(ncol, nlay, seed=6)
    ks = spectral.synthetic_kdist_sw(ngpt)
    k_sw = api.ty_gas_optics_rrtmgp(ctx); k_sw.load(ks)
    nets = H.device_nets(ctx, files)
    out = {}
    for tcf in (0, 1):
        ctx.set_flag("nn_tensor_cores", tcf)
        op = api.ty_optical_props_2str(); op.alloc_2str(ncol, nlay, k_sw)
        op.tau.fill_(-777.0); op.ssa.fill_(-777.0)
        toa = torch.empty((ncol, ngpt), device="cuda")
        msg = k_sw.gas_optics(atm["play"], atm["plev"], atm["tlay"], H.gas_concs(atm["gases"]), op, toa, neural_nets=nets)
        assert msg == "", msg
        torch.cuda.synchronize()
        out[tcf] = {"tau": op.tau.cpu().numpy(), "ssa": op.ssa.cpu().numpy()}
    print(f"SW ncol {ncol} nlay {nlay} ngpt {ngpt}:")
    return max(report(k, out[1][k], out[0][k]) for k in out[0])


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, run_lw(7, 60, H.LW_G256, 256))
    worst = max(worst, run_lw(50, 60, H.LW_G256, 256))
    worst = max(worst, run_lw(33, 137, H.LW_G256, 256, flip=True))
    worst = max(worst, run_lw(301, 33, H.LW_G256, 256))
    worst = max(worst, run_sw(7, 60, H.SW_G224, 224))
    worst = max(worst, run_sw(45, 60, H.SW_G224, 224))
    worst = max(worst, run_sw(301, 33, H.SW_G224, 224))
    worst = max(worst, run_sw(2000, 137, H.SW_G224, 224))
    worst = max(worst, run_lw(2000, 137, H.LW_G256, 256))
    print("worst", worst)
    sys.exit(0 if worst < 5e-4 else 1)
