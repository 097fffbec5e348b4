"""(needs a library built with make EXTRA=-DRRNN_EXPERIMENT_SW_WIDE)  A/B of the wide SW solver (sw_solver_v7, four g-points per lane; context flag solver_wide_sw) against sw_solver_v6 on the same
inputs: max |flux difference| relative to the largest flux, several shapes / orientations, with and without diffuse incident flux."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib

ctx = api.default_context(0)
P = api._ptr
g = torch.Generator(device="cuda").manual_seed(5)
mk = lambda *s: torch.rand(*s, device="cuda", generator=g)
worst = 0.0
for (ncol, nlay, G) in ((301, 137, 224), (77, 61, 224), (1000, 16, 112), (33, 9, 224), (500, 60, 132), (2500, 60, 256)):
    tau = mk(ncol, nlay, G) * 0.5; tau[:, ::7, ::5] *= 1e-5
    ssa = mk(ncol, nlay, G); ssa[:, ::3, ::4] *= 1e-3
    mu0 = mk(ncol) * 0.9 + 0.1; inc = mk(ncol, G); dif = mk(ncol, G) * 0.1; alb = mk(ncol, G) * 0.5; alb2 = mk(ncol, G) * 0.5
    for top in (1, 0):
        for with_dif in (0, 1):
            for fast in (1, 0):
                ctx.set_flag("sw_fast_math", fast)
                out = []
                for wide in (0, 1):
                    ctx.set_flag("solver_wide_sw", wide)
                    fl = [torch.zeros(ncol, nlay + 1, device="cuda") for _ in range(3)]
                    _lib.check(_lib.lib().rrnn_sw_solver_2stream(ctx.h, G, nlay, ncol, top, P(inc), P(dif) if with_dif else None, P(tau), P(ssa), None, P(mu0),
                                                                 P(alb), P(alb2), P(fl[0]), P(fl[1]), P(fl[2])))
                    torch.cuda.synchronize()
                    out.append(fl)
                d = max(float((out[0][i] - out[1][i]).abs().max() / out[0][i].abs().max()) for i in range(3))
                worst = max(worst, d)
                print(json.dumps({"ncol": ncol, "nlay": nlay, "ngpt": G, "top_at_1": top, "inc_flux_dif": with_dif, "sw_fast_math": fast, "max_rel_diff_wide_vs_v6": d}))
ctx.set_flag("sw_fast_math", 1)
print("WORST", worst)
assert worst < 3e-6
