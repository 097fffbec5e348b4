"""A/B of the wide LW solver (lw_solver_v7, four g-points per lane; context flag solver_wide) against lw_solver_v6 on the same inputs:
max |flux difference| relative to the largest flux, materialised sources, several shapes / orientations / angle counts."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib

ctx = api.default_context(0)
P = api._ptr
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda *s: torch.rand(*s, device="cuda", generator=g)
worst = 0.0
for (ncol, nlay, G) in ((301, 137, 256), (77, 61, 256), (1000, 16, 128), (33, 9, 256), (500, 60, 132)):
    tau = mk(ncol, nlay, G) * 2.0; tau[:, ::7, ::5] *= 1e-5
    lay = mk(ncol, nlay, G) + 0.5; lev = mk(ncol, nlay + 1, G) + 0.5; emis = mk(ncol, G) * 0.2 + 0.8; ss = mk(ncol, G) + 0.5; inc = mk(ncol, G) * 0.1
    for top in (1, 0):
        for nmus in (1, 3):
            Ds = np.array([1.66, 1.2, 2.5][:nmus], np.float32); w = np.array([0.5, 0.3, 0.2][:nmus], np.float32)
            out = []
            for wide in (0, 1):
                ctx.set_flag("solver_wide", wide)
                fu = torch.zeros(ncol, nlay + 1, device="cuda"); fd = torch.zeros(ncol, nlay + 1, device="cuda")
                _lib.check(_lib.lib().rrnn_lw_solver_noscat(ctx.h, G, nlay, ncol, top, nmus, Ds.ctypes.data_as(_lib.c_float_p), w.ctypes.data_as(_lib.c_float_p),
                                                            P(inc), P(tau), P(lay), P(lev), P(emis), P(ss), P(fu), P(fd)))
                torch.cuda.synchronize()
                out.append((fu, fd))
            d = max(float((out[0][i] - out[1][i]).abs().max() / out[0][i].abs().max()) for i in (0, 1))
            worst = max(worst, d)
            print(json.dumps({"ncol": ncol, "nlay": nlay, "ngpt": G, "top_at_1": top, "nmus": nmus, "max_rel_diff_wide_vs_v6": d}))
print("WORST", worst)
assert worst < 2e-6
