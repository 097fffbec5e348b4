"""GPU tests at BASELINE.json's full sizes, through properties that do not need the oracle at that size.

configs[3] (1 000 000 columns x 137 layers, clear-sky LW+SW) and configs[2] (100 000 x 60 all-sky) are run through the
C ABI on inputs tiled from a few thousand distinct columns.  Every column is an independent problem
(SURVEY.md 8e), so
  * every replica of a base column must give the SAME fluxes bit for bit, wherever it falls in a chunk, a 128-row MLP
    tile or a solver cluster (the reference's own invariance tests: column subsets and block-size independence,
    tests/clear_sky_regression.F90:225-300);
  * the first n0 columns must equal a separate n0-column call bit for bit;
  * a sample of the base columns is checked against the oracle within the stated flux tolerance;
  * boundary conditions hold at every column: no LW flux down at the top, SW flux down at the top = mu0 x sum of the
    solar source = the direct flux there, direct <= total, flux_up / flux_dn of night columns zero, everything finite.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _tile(t, ncol):
    reps = -(-ncol // t.shape[0])
    return t.repeat((reps,) + (1,) * (t.dim() - 1))[:ncol].contiguous()


def _replicas_identical(full, n0):
    """full[(k*n0 + j)] == full[j] for every replica k (bit-exact)."""
    torch = _torch()
    ncol = full.shape[0]
    K = ncol // n0
    base = full[:n0]
    body = full[:K * n0].view(K, n0, *full.shape[1:])
    ok = bool((body == base.unsqueeze(0)).all())
    rem = ncol - K * n0
    if rem:
        ok = ok and torch.equal(full[K * n0:], base[:rem])
    return ok


def test_clear_sky_lw_sw_at_1M_columns_137_layers(gpu_ctx):
    import oracle as O
    from rte_rrtmgp_nn_b200 import api, spectral, synth
    torch = _torch()
    ncol, nlay, n0 = 1_000_000, 137, 2048
    nlev = nlay + 1
    base = synth.make_atmosphere(n0, nlay, seed=4242)
    base["mu0"][::11] = -0.2                                   # night columns (rrtmgp_rfmip_sw.F90:458-463)
    dev = torch.device("cuda", gpu_ctx.device)
    b = {k: torch.from_numpy(np.ascontiguousarray(base[k], np.float32)).to(dev)
         for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0")}
    big = {k: _tile(v, ncol) for k, v in b.items()}

    def concs(tile_to=None):
        gc = api.ty_gas_concs()
        for k, v in base["gases"].items():
            if np.ndim(v) == 2:
                t = torch.from_numpy(np.ascontiguousarray(v, np.float32)).to(dev)
                gc.set_vmr(k, _tile(t, tile_to) if tile_to else t)
            else:
                gc.set_vmr(k, float(v))
        return gc

    kdl, kds = spectral.synthetic_kdist_lw(256), spectral.synthetic_kdist_sw(224)
    k_lw = api.ty_gas_optics_rrtmgp(gpu_ctx); assert k_lw.load(kdl) == ""
    k_sw = api.ty_gas_optics_rrtmgp(gpu_ctx); assert k_sw.load(kds) == ""
    nl, ns = H.device_nets(gpu_ctx, H.LW_G256), H.device_nets(gpu_ctx, H.SW_G224)

    def run(d, gc, n):
        fl = {k: torch.full((n, nlev), float("nan"), device=dev) for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir")}
        assert api.lw_fluxes(k_lw, nl, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gc, fl["lw_up"], fl["lw_dn"],
                             tlev=d["tlev"], top_at_1=True, n_gauss_angles=1) in ("", None)
        assert api.sw_fluxes(k_sw, ns, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gc, fl["sw_up"], fl["sw_dn"],
                             fl["sw_dir"], top_at_1=True) in ("", None)
        gpu_ctx.synchronize()
        return fl

    full = run(big, concs(ncol), ncol)
    small = run(b, concs(), n0)
    for k, v in full.items():
        assert bool(torch.isfinite(v).all()), k
        assert _replicas_identical(v, n0), f"{k}: replicas of one column differ across the 1M-column call"
        assert torch.equal(v[:n0], small[k]), f"{k}: the first {n0} columns differ from a {n0}-column call"
    # boundary conditions at every column
    night = big["mu0"] <= 0
    assert bool((full["lw_dn"][:, 0] == 0).all())
    assert bool((full["lw_up"] > 0).all()) and bool((full["lw_dn"][:, 1:] > 0).all())
    for k in ("sw_up", "sw_dn"):   # the drivers zero flux_up / flux_dn of night columns only (rrtmgp_rfmip_sw.F90:458-463)
        assert bool((full[k][night] == 0).all()), k
    day = ~night
    tsi = float(np.asarray(kds["solar_source"], np.float64).sum())
    assert torch.equal(full["sw_dn"][day][:, 0], full["sw_dir"][day][:, 0])
    assert torch.allclose(full["sw_dn"][day][:, 0], big["mu0"][day] * tsi, rtol=2e-5, atol=0)
    assert bool((full["sw_dir"][day] <= full["sw_dn"][day] + 1e-3).all()) and bool((full["sw_up"] >= 0).all())
    # a sample of the base columns against the oracle (the 1M-column results are these, replicated)
    idx = np.arange(0, n0, n0 // 24)[:24]
    sub = {k: np.ascontiguousarray(v[idx]) for k, v in base.items() if isinstance(v, np.ndarray)}
    gases = {k: (np.ascontiguousarray(v[idx]) if np.ndim(v) == 2 else v) for k, v in base["gases"].items()}
    got = {k: v[torch.from_numpy(idx).to(dev)].cpu().numpy() for k, v in full.items()}
    onl, ons = H.oracle_nets(H.LW_G256), H.oracle_nets(H.SW_G224)
    ref = O.gas_optics_lw(kdl, onl, sub["play"], sub["plev"], sub["tlay"], sub["tsfc"], gases, tlev=sub["tlev"])
    rup, rdn = O.rte_lw(kdl, True, ref["tau"], ref["lay_source"], ref["lev_source"], ref["sfc_source"],
                        np.repeat(sub["sfc_emis"][:, None], 16, 1))
    assert np.abs(got["lw_up"] - rup).max() <= H.FLUX_TOL and np.abs(got["lw_dn"] - rdn).max() <= H.FLUX_TOL
    mu0 = sub["mu0"]; mu0e = np.where(mu0 > 0, mu0, 1.0).astype(np.float32)
    alb = np.repeat(sub["sfc_alb"][:, None], 224, 1)
    outs = []
    for fast in (False, "f64"):
        r = O.gas_optics_sw(kds, ons, sub["play"], sub["plev"], sub["tlay"], gases, fast=fast)
        f = [np.array(a) for a in O.rte_sw(True, mu0e, r["toa_src"], alb, alb, r["tau"], r["ssa"], r["g"], fast=fast)]
        for a in f[:2]:
            a[mu0 <= 0] = 0
        outs.append(f)
    for k, w32, w64 in zip(("sw_up", "sw_dn", "sw_dir"), outs[0], outs[1]):
        H.assert_within_reference_noise(got[k], w32, w64, H.FLUX_TOL, f"1M x 137 {k}")


def test_all_sky_lw_at_100k_columns_60_layers(gpu_ctx):
    """configs[2] at full size on the stage API: gas optics -> LUT cloud optics -> increment -> rte_lw, replicas bit-identical
    and equal to a small call; clear columns (every third one, rrtmgp_allsky.F90:333-350) equal the clear-sky solution."""
    import os
    from rte_rrtmgp_nn_b200 import api, spectral, synth
    torch = _torch()
    ncol, nlay, n0 = 100_000, 60, 1536
    dev = torch.device("cuda", gpu_ctx.device)
    base = synth.make_atmosphere(n0, nlay, seed=777)
    lut_path = os.path.join(H.ROOT, "data", "cloud_optics", "rrtmgp-cloud-optics-coeffs-lw.nc")
    kd = spectral.synthetic_kdist_lw(256)
    k_dist = api.ty_gas_optics_rrtmgp(gpu_ctx); assert k_dist.load(kd) == ""
    nets = H.device_nets(gpu_ctx, H.LW_G256)
    cloud_optics = api.ty_cloud_optics(gpu_ctx)
    assert cloud_optics.load(**api.load_cloud_lut_file(lut_path)) == ""

    def solve(n, with_clouds=True):
        t = lambda a: _tile(torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev), n)
        gc = api.ty_gas_concs()
        for k, v in base["gases"].items():
            gc.set_vmr(k, t(v) if np.ndim(v) == 2 else float(v))
        op = api.ty_optical_props_1scl(); assert op.alloc_1scl(n, nlay, k_dist) == ""
        src = api.ty_source_func_lw(); assert src.alloc(n, nlay, k_dist) == ""
        assert k_dist.gas_optics(t(base["play"]), t(base["plev"]), t(base["tlay"]), t(base["tsfc"]), gc, op, src,
                                 tlev=t(base["tlev"]), neural_nets=nets) == ""
        if with_clouds:
            play = t(base["play"])
            col = torch.arange(n, device=dev) % n0
            mask = ((play > 1.0e4) & (play < 9.0e4) & ((col % 3) != 0)[:, None]).float()
            clouds = api.ty_optical_props_1scl(); assert clouds.alloc_1scl(n, nlay, k_dist, by_band=True) == ""
            re_l = torch.full((n, nlay), 10.0, device=dev); re_i = torch.full((n, nlay), 50.0, device=dev)
            assert cloud_optics.cloud_optics(10.0 * mask, 10.0 * mask, re_l, re_i, clouds) == ""
            assert clouds.increment(op) == ""
        fl = api.ty_fluxes_broadband(torch.zeros((n, nlay + 1), device=dev), torch.zeros((n, nlay + 1), device=dev))
        emis = torch.full((n, 16), 0.98, device=dev)
        assert api.rte_lw(op, True, src, emis, fl) == ""
        gpu_ctx.synchronize()
        return fl

    full, small, clear = solve(ncol), solve(n0), solve(n0, with_clouds=False)
    for nm in ("flux_up", "flux_dn"):
        v = getattr(full, nm)
        assert bool(torch.isfinite(v).all())
        assert _replicas_identical(v, n0), nm
        assert torch.equal(v[:n0], getattr(small, nm)), nm
        assert torch.equal(getattr(small, nm)[0::3], getattr(clear, nm)[0::3]), nm      # cloud-free columns
    assert float((small.flux_dn[1::3] - clear.flux_dn[1::3]).abs().max()) > 1.0          # the clouds are there
