"""Debug driver for the tensor-core gas optics: one call, TC path only (RRNN_TC_DEBUG controls the kernel's debug switches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import helpers as H
from rte_rrtmgp_nn_b200 import api, spectral, synth
mode, ncol, nlay = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = api.default_context(0)
ctx.set_flag("nn_tensor_cores", 1)
atm = synth.make_atmosphere(ncol, nlay, seed=5)
if mode == "lw":
    k = api.ty_gas_optics_rrtmgp(ctx); k.load(spectral.synthetic_kdist_lw(256))
    nets = H.device_nets(ctx, H.LW_G256)
    op = api.ty_optical_props_1scl(); op.alloc_1scl(ncol, nlay, k)
    src = api.ty_source_func_lw(); src.alloc(ncol, nlay, k)
    for t in (op.tau, src.lay_source, src.lev_source, src.sfc_source, src.sfc_source_Jac): t.fill_(-777.0)
    print(k.gas_optics(atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], H.gas_concs(atm["gases"]), op, src, tlev=atm["tlev"], neural_nets=nets))
    torch.cuda.synchronize()
    for n, t in (("tau", op.tau), ("lay", src.lay_source), ("lev", src.lev_source), ("sfc", src.sfc_source)):
        a = t.cpu().numpy(); print(n, "untouched", int((a == -777.0).sum()), "of", a.size, "min", np.nanmin(a), "max", np.nanmax(a))
else:
    k = api.ty_gas_optics_rrtmgp(ctx); k.load(spectral.synthetic_kdist_sw(224))
    nets = H.device_nets(ctx, H.SW_G224)
    op = api.ty_optical_props_2str(); op.alloc_2str(ncol, nlay, k)
    op.tau.fill_(-777.0); op.ssa.fill_(-777.0)
    toa = torch.empty((ncol, 224), device="cuda")
    print(k.gas_optics(atm["play"], atm["plev"], atm["tlay"], H.gas_concs(atm["gases"]), op, toa, neural_nets=nets))
    torch.cuda.synchronize()
    for n, t in (("tau", op.tau), ("ssa", op.ssa)):
        a = t.cpu().numpy(); print(n, "untouched", int((a == -777.0).sum()), "of", a.size, "min", np.nanmin(a), "max", np.nanmax(a))
