import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/oracle', '/root/repo/tests']
import numpy as np, torch
import helpers as H, oracle as O
from rte_rrtmgp_nn_b200 import api, spectral, synth
ctx = api.default_context(0)
for (ncol, nlay, seed) in ((200, 137, 3), (300, 60, 4)):
    atm = synth.make_atmosphere(ncol, nlay, seed=seed)
    kd = spectral.synthetic_kdist_lw(256)
    k = api.ty_gas_optics_rrtmgp(ctx); k.load(kd)
    nets = H.device_nets(ctx, H.LW_G256); onets = H.oracle_nets(H.LW_G256)
    ref = O.gas_optics_lw(kd, onets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"], fast="f64")
    emis = np.repeat(atm["sfc_emis"][:, None], 16, 1)
    r64 = O.rte_lw(kd, True, ref["tau"], ref["lay_source"], ref["lev_source"], ref["sfc_source"], emis, fast="f64")
    for fm in (0, 1):
        ctx.set_flag("fast_math", fm)
        up, dn = api.lw_fluxes_host(k, nets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["sfc_emis"], H.gas_concs(atm["gases"]), tlev=atm["tlev"])
        hr = lambda u, d: O.calc_heating_rate(u.astype(np.float32), d.astype(np.float32), atm["plev"])
        dhr = np.abs(hr(up, dn) - hr(r64[0], r64[1]))
        thick = np.abs(np.diff(atm["plev"], axis=1)) >= 500
        print(ncol, nlay, "fast_math", fm, "max|dflux| vs f64:", np.abs(up - r64[0]).max(), np.abs(dn - r64[1]).max(), "max|dHR| thick layers", dhr[thick].max())
ctx.set_flag("fast_math", 0)
