"""Time the SW / LW solver kernels alone (CUDA events on the context's stream) on synthetic optical properties.

    python tools/time_solvers.py [sw|lw|both] [ncol] [nlay] [reps] [flag=value ...]

Prints one JSON line per kernel: ms per launch, columns/s, algorithmic GB/s and fraction of the measured HBM peak.
RRNN_LIB_PATH selects another build of the library (A/B experiments)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "both"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
nlay = int(sys.argv[3]) if len(sys.argv) > 3 else 137
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = api.default_context(0)
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    ctx.set_flag(k, int(v))
P = api._ptr
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda *s: torch.rand(*s, device="cuda", generator=g)
fl = [torch.empty((ncol, nlay + 1), device="cuda") for _ in range(3)]
peak = 6466.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
stream = ctx.torch_stream()


def timeit(fn):
    for _ in range(2):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda.synchronize()
    for i in range(reps):
        ev[i].record(stream)
        fn()
    ev[reps].record(stream)
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


if which in ("sw", "both"):
    G = int(os.environ.get("G_SW", 224))
    tau = mk(ncol, nlay, G) * 0.5; ssa = mk(ncol, nlay, G); mu0 = mk(ncol) * 0.9 + 0.1; inc = mk(ncol, G); alb = mk(ncol, G) * 0.5
    f = lambda: _lib.check(_lib.lib().rrnn_sw_solver_2stream(ctx.h, G, nlay, ncol, 1, P(inc), None, P(tau), P(ssa), None, P(mu0), P(alb), P(alb),
                                                             P(fl[0]), P(fl[1]), P(fl[2])))
    t = timeit(f)
    by = 4 * G * (2 * nlay + 3) + 4 + 12 * (nlay + 1)
    med = t[len(t) // 2]
    print(json.dumps({"kernel": "sw_solver", "ncol": ncol, "nlay": nlay, "ngpt": G, "ms_median": med, "ms_min": t[0], "ms_per_1M_columns": med * 1e6 / ncol,
                      "algorithmic_gb_per_s": by * ncol / med / 1e6, "frac_of_hbm_peak": by * ncol / med / 1e6 / peak,
                      "checksum": [float(x.double().sum()) for x in fl]}))
    del tau, ssa
if which in ("lw", "both"):
    G = int(os.environ.get("G_LW", 256))
    tau = mk(ncol, nlay, G) * 0.5; lay = mk(ncol, nlay, G); lev = mk(ncol, nlay + 1, G); emis = mk(ncol, G); ss = mk(ncol, G)
    Ds = np.array([1.66], np.float32); w = np.array([0.5], np.float32)
    f = lambda: _lib.check(_lib.lib().rrnn_lw_solver_noscat(ctx.h, G, nlay, ncol, 1, 1, Ds.ctypes.data_as(_lib.c_float_p), w.ctypes.data_as(_lib.c_float_p),
                                                            None, P(tau), P(lay), P(lev), P(emis), P(ss), P(fl[0]), P(fl[1])))
    t = timeit(f)
    by = 4 * G * (3 * nlay + 3) + 8 * (nlay + 1)
    med = t[len(t) // 2]
    print(json.dumps({"kernel": "lw_solver (materialised sources)", "ncol": ncol, "nlay": nlay, "ngpt": G, "ms_median": med, "ms_min": t[0],
                      "ms_per_1M_columns": med * 1e6 / ncol, "algorithmic_gb_per_s": by * ncol / med / 1e6,
                      "frac_of_hbm_peak": by * ncol / med / 1e6 / peak, "checksum": [float(x.double().sum()) for x in fl[:2]]}))
