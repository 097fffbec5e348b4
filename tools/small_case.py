"""Latency of one LW + SW pass at the RFMIP shape (1800 columns x 60 layers; BASELINE configs[0], [1]) through the
device-pointer drivers: wall clock per pass (host launch overhead included) and device time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
from rte_rrtmgp_nn_b200 import api, spectral
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1800
nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ctx = api.default_context(0)
atm = bench.make_inputs(ncol, nlay)
k_lw = api.ty_gas_optics_rrtmgp(ctx); k_lw.load(spectral.synthetic_kdist_lw(256))
k_sw = api.ty_gas_optics_rrtmgp(ctx); k_sw.load(spectral.synthetic_kdist_sw(224))
nl = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(bench.NN_DIR, f)) for f in bench.LW_FILES]
ns = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(bench.NN_DIR, f)) for f in bench.SW_FILES]
d = {k: torch.from_numpy(atm[k]).cuda() for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0")}
gc = api.ty_gas_concs()
for k, v in atm["gases"].items():
    gc.set_vmr(k, torch.from_numpy(v).cuda() if np.ndim(v) == 2 else float(v))
fl = [torch.empty((ncol, nlay + 1), device="cuda") for _ in range(5)]
def one():
    api.lw_fluxes(k_lw, nl, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gc, fl[0], fl[1], tlev=d["tlev"])
    api.sw_fluxes(k_sw, ns, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gc, fl[2], fl[3], fl[4])
for _ in range(5): one()
torch.cuda.synchronize()
n = 50
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): one()
e1.record(); t_launch = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"{ncol} x {nlay}: wall {1e3*t_all/n:.3f} ms/pass (host enqueue {1e3*t_launch/n:.3f} ms), device {e0.elapsed_time(e1)/n:.3f} ms/pass, "
      f"{ncol/(t_all/n):.0f} columns/s")
