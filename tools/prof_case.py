"""Small fixed case for ncu: one LW+SW pass over NCOL columns x 137 layers (device-resident inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
from rte_rrtmgp_nn_b200 import api, spectral

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 137
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fast = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = api.default_context(0)
ctx.set_flag("fast_math", fast)
for kv in sys.argv[5:]:      # further context flags, e.g. solver_wide=1
    k_, v_ = kv.split("=")
    ctx.set_flag(k_, int(v_))
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
def one_pass():
    api.lw_fluxes(k_lw, nl, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gc, fl[0], fl[1], tlev=d["tlev"])
    api.sw_fluxes(k_sw, ns, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gc, fl[2], fl[3], fl[4])
if reps > 2:       # (ncu captures run with reps <= 2 and profile every launch)
    one_pass()     # untimed: workspace / scratch allocation, first-launch overheads
    torch.cuda.synchronize()
ctx.profile(True)
for _ in range(reps):
    api.lw_fluxes(k_lw, nl, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gc, fl[0], fl[1], tlev=d["tlev"])
    api.sw_fluxes(k_sw, ns, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gc, fl[2], fl[3], fl[4])
torch.cuda.synchronize()
print({k: (round(v[0] / max(1, v[1]), 3), v[1]) for k, v in ctx.profile_read().items()})
print("checksum", float(fl[0].sum()), float(fl[3].sum()))
