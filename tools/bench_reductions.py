"""Timing of the reductions either side of the solvers (fluxes_ext.cu: by-band sums, net flux, optimal angles) against
their algorithmic bytes, and of rte_sw with a ty_fluxes_byband beside the plain broadband call:
python tools/bench_reductions.py [ncol nlay]  -> one JSON line per case"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
from rte_rrtmgp_nn_b200 import api, _lib, spectral
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = int(sys.argv[2]) if len(sys.argv) > 2 else 60
PEAK = 6466.8
ctx = api.default_context(0)
gen = torch.Generator(device="cuda").manual_seed(1)
mk = lambda *s: torch.rand(*s, device="cuda", generator=gen)
P = api._ptr
lib = _lib.lib()


def timed(f, reps=20, trials=3):
    """median over `trials` of the mean over `reps` back-to-back launches (the arrays are larger than the 126 MB L2)"""
    for _ in range(3): f()
    torch.cuda.synchronize()
    out = []
    for _ in range(trials):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()       # the default context launches on the legacy default stream = torch's default stream
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    return sorted(out)[len(out) // 2]


def line(case, ms, gbytes, **kw):
    print(json.dumps(dict(case=case, ncol=ncol, nlay=L, ms=round(ms, 4), columns_per_s=round(ncol / ms * 1e3),
                          algorithmic_gb_per_s=round(gbytes / ms * 1e3, 1), frac_of_hbm_peak=round(gbytes / ms * 1e3 / PEAK, 3), **kw)))


for name, kd in (("sw g224", spectral.synthetic_kdist_sw(224)), ("lw g256", spectral.synthetic_kdist_lw(256))):
    G, B = int(kd["ngpt"]), int(kd["nbnd"])
    h = api._kdist_handle(ctx, dict(kd, optimal_angle_fit=np.tile(np.array([[0.3, 1.6]], np.float32), (B, 1))))
    up, dn = mk(ncol, L + 1, G), mk(ncol, L + 1, G)
    bnd = torch.empty((ncol, L + 1, B), device="cuda")
    rows = ncol * (L + 1)
    line(f"sum_byband {name}", timed(lambda: _lib.check(lib.rrnn_sum_byband(ctx.h, h.h, L + 1, ncol, P(up), P(bnd)))), 4e-9 * rows * (G + B))
    line(f"net_byband {name}", timed(lambda: _lib.check(lib.rrnn_net_byband(ctx.h, h.h, L + 1, ncol, P(dn), P(up), P(bnd)))), 4e-9 * rows * (2 * G + B))
    net = torch.empty_like(up)
    line(f"net_flux {name} (g-point arrays)", timed(lambda: _lib.check(lib.rrnn_net_flux(ctx.h, up.numel(), P(dn), P(up), P(net)))), 12e-9 * rows * G)
    tau = mk(ncol, L, G) * 0.1
    ang = torch.empty((ncol, G), device="cuda")
    line(f"compute_optimal_angles {name}", timed(lambda: _lib.check(lib.rrnn_compute_optimal_angles(ctx.h, h.h, L, ncol, P(tau), P(ang)))),
         4e-9 * ncol * G * (L + 1))
    del up, dn, bnd, net, tau, ang

# rte_sw: broadband only (tuned kernel) / with flux_net / with by-band fluxes (general kernel + reductions)
G = 224
k_dist = api.ty_gas_optics_rrtmgp(ctx); k_dist.load(spectral.synthetic_kdist_sw(G))
atmos = api.ty_optical_props_2str(); atmos.alloc_2str(ncol, L, k_dist)
atmos.tau.copy_(mk(ncol, L, G) * 0.5); atmos.ssa.copy_(mk(ncol, L, G)); atmos.g_is_zero = True
mu0 = mk(ncol) * 0.9 + 0.1; inc = mk(ncol, G); alb = mk(ncol, G) * 0.5
z = lambda *s: torch.empty(s, device="cuda")
gb_in = 8e-9 * ncol * L * G
plain = api.ty_fluxes_broadband(z(ncol, L + 1), z(ncol, L + 1), None, z(ncol, L + 1))
line("rte_sw broadband (sw_solver_v5)", timed(lambda: api.rte_sw(atmos, True, mu0, inc, alb, alb, plain)), gb_in)
wnet = api.ty_fluxes_broadband(z(ncol, L + 1), z(ncol, L + 1), z(ncol, L + 1), z(ncol, L + 1))
line("rte_sw broadband + flux_net", timed(lambda: api.rte_sw(atmos, True, mu0, inc, alb, alb, wnet)), gb_in)
bb = api.ty_fluxes_byband(z(ncol, L + 1), z(ncol, L + 1), z(ncol, L + 1), z(ncol, L + 1), z(ncol, L + 1, 14), z(ncol, L + 1, 14),
                          z(ncol, L + 1, 14), z(ncol, L + 1, 14))
line("rte_sw with ty_fluxes_byband (by-band sums inside the tuned solver; was: general kernel + g-point temporaries + reductions)",
     timed(lambda: api.rte_sw(atmos, True, mu0, inc, alb, alb, bb), reps=5, trials=2), gb_in + 3 * 2 * 4e-9 * ncol * (L + 1) * G)
