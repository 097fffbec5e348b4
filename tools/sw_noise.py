"""How far is each SW solver variant from the fp64 evaluation of the same equations on the SAME fp32 tau/ssa?
(separates solver rounding noise from gas-optics differences; run under gpurun)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import helpers as H, oracle as O
from rte_rrtmgp_nn_b200 import api, spectral, synth, _lib
ctx = api.default_context(0)
P = api._ptr
def stats(a, b):
    d = np.abs(np.asarray(a, np.float64) - b)
    return "max %.3e rms %.3e" % (d.max(), np.sqrt((d ** 2).mean()))
for (ncol, nlay, flip, seed) in ((5, 137, True, 2), (200, 137, False, 7), (200, 60, False, 8)):
    kd = spectral.synthetic_kdist_sw(ngpt=224)
    atm = synth.make_atmosphere(ncol, nlay, seed=seed)
    if flip: atm = synth.flip_vertical(atm)
    nets = H.oracle_nets(H.SW_G224)
    ref = O.gas_optics_sw(kd, nets, atm["play"], atm["plev"], atm["tlay"], atm["gases"])
    alb = np.repeat(atm["sfc_alb"][:, None], 224, 1)
    top = atm["top_at_1"]
    u32, d32, r32 = O.rte_sw(top, atm["mu0"], ref["toa_src"], alb, alb, ref["tau"], ref["ssa"], ref["g"])
    u64, d64, r64 = O.rte_sw(top, atm["mu0"], ref["toa_src"], alb, alb, ref["tau"].astype(np.float64), ref["ssa"].astype(np.float64), ref["g"].astype(np.float64), fast="f64")
    print(f"== ncol {ncol} nlay {nlay} flip {flip}: oracle32 vs fp64  up {stats(u32,u64)}  dn {stats(d32,d64)}")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(inc=ref["toa_src"], tau=ref["tau"], ssa=ref["ssa"], mu0=atm["mu0"], alb=alb).items()}
    mk = lambda: torch.empty((ncol, nlay + 1), device="cuda")
    for variant, name in ((0, "v5 tma packed"), (2, "v4 packed"), (1, "v3 scalar")):
        for fast in (0, 1):
            ctx.set_flag("solver_variant", variant); ctx.set_flag("fast_math", fast)
            up, dn, dr = mk(), mk(), mk()
            _lib.check(_lib.lib().rrnn_sw_solver_2stream(ctx.h, 224, nlay, ncol, int(top), P(d["inc"]), None, P(d["tau"]), P(d["ssa"]), None, P(d["mu0"]), P(d["alb"]), P(d["alb"]), P(up), P(dn), P(dr)))
            torch.cuda.synchronize()
            print(f"   {name:14s} fast={fast}: up {stats(up.cpu().numpy(),u64)}  dn {stats(dn.cpu().numpy(),d64)}  dir {stats(dr.cpu().numpy(),r64)}")
    ctx.set_flag("solver_variant", 0); ctx.set_flag("fast_math", 0)
