"""Timing of the general LW kernel (re-scaled scattering / Jacobian / g-point fluxes, rrnn_lw_solver_noscat_ext) beside the
tuned no-scattering kernel and the CPU oracle: python tools/bench_lw_ext.py [ncol nlay]  -> one JSON line per case"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = int(sys.argv[2]) if len(sys.argv) > 2 else 60
G = 256
ctx = api.default_context(0)
g_ = torch.Generator(device="cuda").manual_seed(1)
mk = lambda *s: torch.rand(*s, device="cuda", generator=g_)
tau = mk(ncol, L, G) * 0.5; ssa = mk(ncol, L, G) * 0.9; asy = mk(ncol, L, G) * 0.8
lay = mk(ncol, L, G) + 1; lev = mk(ncol, L + 1, G) + 1; emis = mk(ncol, G) * 0.2 + 0.8; ss = mk(ncol, G) + 1; sj = ss * 0.01
up = torch.empty((ncol, L + 1), device="cuda"); dn = torch.empty_like(up); jac = torch.empty_like(up)
Ds = np.array([1.66], np.float32); w = np.array([0.5], np.float32)
fp = lambda a: a.ctypes.data_as(_lib.c_float_p)
P = api._ptr
lib = _lib.lib()
def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.ExternalStream(ctx.stream_ptr()) if hasattr(ctx, "stream_ptr") else None)
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
cases = {
    "rescaled (ssa, g)": (lambda: _lib.check(lib.rrnn_lw_solver_noscat_ext(ctx.h, G, L, ncol, 1, 1, fp(Ds), fp(w), None, None, P(tau), P(ssa), P(asy), P(lay), P(lev), P(emis), P(ss), None, P(up), P(dn), None, None, None)), 5),
    "rescaled + Jacobian": (lambda: _lib.check(lib.rrnn_lw_solver_noscat_ext(ctx.h, G, L, ncol, 1, 1, fp(Ds), fp(w), None, None, P(tau), P(ssa), P(asy), P(lay), P(lev), P(emis), P(ss), P(sj), P(up), P(dn), P(jac), None, None)), 5),
    "no scattering, general kernel": (lambda: _lib.check(lib.rrnn_lw_solver_noscat_ext(ctx.h, G, L, ncol, 1, 1, fp(Ds), fp(w), None, None, P(tau), None, None, P(lay), P(lev), P(emis), P(ss), None, P(up), P(dn), None, None, None)), 3),
    "no scattering, tuned kernel (lw_solver_v5)": (lambda: _lib.check(lib.rrnn_lw_solver_noscat(ctx.h, G, L, ncol, 1, 1, fp(Ds), fp(w), None, P(tau), P(lay), P(lev), P(emis), P(ss), P(up), P(dn))), 3),
}
for name, (f, narr) in cases.items():
    ms = timed(f)
    gb = narr * 4.0 * G * L * ncol / 1e9
    print(json.dumps({"case": name, "ncol": ncol, "nlay": L, "ngpt": G, "ms": round(ms, 3), "columns_per_s": round(ncol / ms * 1e3),
                      "algorithmic_gb_per_s": round(gb / ms * 1e3, 1), "frac_of_hbm_peak_6466.8": round(gb / ms * 1e3 / 6466.8, 3)}))
# the CPU oracle beside it, on a bounded sample
import oracle as O
n = min(ncol, 2048)
a = [t[:n].cpu().numpy() for t in (tau, lay, lev, emis, ss, ssa, asy)]
t0 = time.perf_counter()
O.lw_solver_noscat_GaussQuad_ext(True, 1, a[0], a[1], a[2], a[3], a[4], ssa=a[5], g=a[6], fast=True)
dt = time.perf_counter() - t0
print(json.dumps({"case": "CPU oracle (-O3 OpenMP), rescaled", "ncol": n, "columns_per_s": round(n / dt), "cores": O.num_threads()}))
