"""One SW (or LW) solver launch on synthetic optical properties for ncu: python tools/prof_sw.py sw|lw ncol nlay scratch_mb [warps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib
which, ncol, nlay, mb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = api.default_context(0)
ctx.set_flag("solver_scratch_mb", mb)
if len(sys.argv) > 5:
    ctx.set_flag("solver_warps", int(sys.argv[5]))
P = api._ptr
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda *s: torch.rand(*s, device="cuda", generator=g)
fl = [torch.empty((ncol, nlay + 1), device="cuda") for _ in range(3)]
if which == "sw":
    G = 224
    tau = mk(ncol, nlay, G) * 0.5; ssa = mk(ncol, nlay, G); mu0 = mk(ncol) * 0.9 + 0.1; inc = mk(ncol, G); alb = mk(ncol, G) * 0.5
    for _ in range(3):
        _lib.check(_lib.lib().rrnn_sw_solver_2stream(ctx.h, G, nlay, ncol, 1, P(inc), None, P(tau), P(ssa), None, P(mu0), P(alb), P(alb), P(fl[0]), P(fl[1]), P(fl[2])))
else:
    G = 256
    tau = mk(ncol, nlay, G) * 0.5; lay = mk(ncol, nlay, G); lev = mk(ncol, nlay + 1, G); emis = mk(ncol, G); ss = mk(ncol, G)
    Ds = np.array([1.66], np.float32); w = np.array([0.5], np.float32)
    for _ in range(3):
        _lib.check(_lib.lib().rrnn_lw_solver_noscat(ctx.h, G, nlay, ncol, 1, 1, Ds.ctypes.data_as(_lib.c_float_p), w.ctypes.data_as(_lib.c_float_p), None, P(tau), P(lay), P(lev), P(emis), P(ss), P(fl[0]), P(fl[1])))
torch.cuda.synchronize()
print("checksum", float(fl[0].sum()), float(fl[1].sum()))
