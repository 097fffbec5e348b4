import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import helpers as H
import test_parity_gpu as T
from rte_rrtmgp_nn_b200 import api, _lib
ctx = api.default_context(0)
P = api._ptr
L_ = _lib.lib()
fp = lambda a: a.ctypes.data_as(_lib.c_float_p)
for (files, G, nlay, ncol, flip, nang, compat) in [(H.LW_G256, 256, 137, 5, True, 1, 1), (H.LW_G256, 256, 64, 9, True, 2, 0), (H.LW_G256, 256, 64, 9, True, 1, 1), (H.LW_G256, 256, 60, 37, False, 1, 1)]:
    kd, atm, k_dist, _, dnets = T._lw_setup(ctx, files, G, ncol, nlay, seed=5, flip=flip)
    op, src = T._run_lw_gas_optics(ctx, k_dist, dnets, atm)
    gc = H.gas_concs(atm["gases"])
    tau = torch.empty((ncol, nlay, G), device="cuda"); pf = torch.empty_like(tau)
    bl = torch.empty((ncol, nlay, 16), device="cuda"); bv = torch.empty((ncol, nlay + 1, 16), device="cuda")
    ss = torch.empty((ncol, G), device="cuda"); sj = torch.empty_like(ss)
    err = k_dist.gas_optics_compact(atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], gc, tau, pf, bl, bv, ss, sj, tlev=atm["tlev"], neural_nets=dnets)
    assert err == "", err
    emis = torch.from_numpy(np.repeat(atm["sfc_emis"][:, None], G, 1).astype(np.float32)).cuda()
    Ds = {1: [1.66], 2: [1.18350343, 2.81649655]}[nang]; ws = {1: [0.5], 2: [0.3180413817, 0.1819586183]}[nang]
    Ds = np.array(Ds, np.float32); ws = np.array(ws, np.float32)
    ctx.set_flag("lw_source_bug_compat", compat)
    for wide in (0, 1):
        ctx.set_flag("solver_wide", wide)
        f = [torch.zeros((ncol, nlay + 1), device="cuda") for _ in range(4)]
        _lib.check(L_.rrnn_lw_solver_noscat(ctx.h, G, nlay, ncol, int(atm["top_at_1"]), nang, fp(Ds), fp(ws), None, P(op.tau), P(src.lay_source), P(src.lev_source), P(emis), P(src.sfc_source), P(f[0]), P(f[1])))
        _lib.check(L_.rrnn_lw_solver_noscat_compact(ctx.h, k_dist._kd.h, nlay, ncol, int(atm["top_at_1"]), nang, fp(Ds), fp(ws), P(tau), P(pf), P(bl), P(bv), P(emis), P(ss), P(f[2]), P(f[3])))
        torch.cuda.synchronize()
        bu = (f[0] != f[2]).nonzero(); bd = (f[1] != f[3]).nonzero()
        print("L", nlay, "ncol", ncol, "top", int(atm["top_at_1"]), "nang", nang, "compat", compat, "wide", wide, "| up diffs", len(bu), bu[:6].tolist(), "| dn diffs", len(bd), bd[:6].tolist())
    ctx.set_flag("lw_source_bug_compat", 1)
