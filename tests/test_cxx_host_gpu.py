"""Compiled host code on the C ABI: examples/cxx/rfmip_driver.cpp (the reference drivers' block loops written on the C++
mirror include/rrnn.hpp) is built with g++, run on the RFMIP columns, and must give the fluxes of the Python host mirror and
of the oracle."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H

CXX_DIR = os.path.join(H.ROOT, "examples", "cxx")
EXE = os.path.join(CXX_DIR, "rfmip_driver")
API_CHECK = os.path.join(CXX_DIR, "api_check")


def _build():
    subprocess.check_call(["make", "-C", CXX_DIR, "-s"])
    assert os.path.exists(EXE) and os.path.exists(API_CHECK)


def _write_case(d, band, atm, kd, block_size, n_quad_angles=1):
    ncol, nlay = atm["play"].shape
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    lines = [f"ncol {ncol}", f"nlay {nlay}", f"block_size {block_size}", f"nbnd {kd['nbnd']}", f"ngpt {kd['ngpt']}",
             f"top_at_1 {int(atm['top_at_1'])}", f"n_quad_angles {n_quad_angles}", f"temp_ref_min {kd.get('temp_ref_min', 0.0)}",
             f"totplnk_delta {kd.get('totplnk_delta', 1.0)}"]
    np.ascontiguousarray(kd["band_lims_gpt"], np.int32).tofile(os.path.join(d, "band_lims.i32"))
    if band == "lw":
        tot = f32(kd["totplnk"]); lines.append(f"ntemp {tot.shape[1]}"); tot.tofile(os.path.join(d, "totplnk.f32"))
        names = ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis")
    else:
        f32(kd["solar_source"]).tofile(os.path.join(d, "solar_source.f32"))
        names = ("play", "plev", "tlay", "mu0", "sfc_alb", "tsi")
        f32(atm["usecol"].astype(np.float32)).tofile(os.path.join(d, "usecol.f32"))
    for k in names:
        f32(atm[k]).tofile(os.path.join(d, k + ".f32"))
    for k, v in atm["gases"].items():
        v = np.asarray(v)
        if v.ndim == 2 and np.ptp(v) == 0:      # well-mixed: a scalar, as the drivers' gas_conc_array holds it
            lines.append(f"gas {k} {float(v.flat[0])!r}")
        elif v.ndim == 2:
            lines.append(f"gasfield {k}"); f32(v).tofile(os.path.join(d, f"gas_{k}.f32"))
        else:
            lines.append(f"gas {k} {float(v)!r}")
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def test_cxx_driver_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback in the compiled host path either: without a CUDA device the driver stops with the library's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _build()
    (tmp_path / "meta.txt").write_text("ncol 1\nnlay 1\nblock_size 1\nnbnd 1\nngpt 1\ntop_at_1 1\n")
    r = subprocess.run([EXE, "lw", str(tmp_path), "a.nc", "b.nc"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device available (this library has no CPU fallback)" in r.stderr
    r = subprocess.run([API_CHECK], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device available (this library has no CPU fallback)" in r.stderr


@pytest.mark.gpu
def test_cxx_mirror_self_check(gpu_ctx):
    """examples/cxx/api_check.cpp: by-band / net fluxes, optimal angles, solar variability and the reference's error strings
    through the C++ mirror, against exact expectations (small integers, a transparent column, the quiet-sun offsets)."""
    _build()
    r = subprocess.run([API_CHECK], capture_output=True, text=True)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("band", ["lw", "sw"])
def test_cxx_rfmip_block_loop_matches_python_mirror_and_oracle(gpu_ctx, tmp_path, band):
    import oracle as O
    from rte_rrtmgp_nn_b200 import api, rfmip, spectral
    _build()
    cols = np.r_[0:24, 700:712, 1790:1800]           # 46 columns of three experiments; blocks of 8 -> ragged last block
    atm = rfmip.load(columns=cols)
    if band == "sw":
        lit = np.flatnonzero(atm["usecol"])
        night, day = int(lit[0]), int(lit[1])
        atm["usecol"][night] = False; atm["mu0"][night] = 1.0   # one more night column, next to a sunlit one
    kd = spectral.synthetic_kdist_lw(256) if band == "lw" else spectral.synthetic_kdist_sw(224)
    files = H.LW_G256 if band == "lw" else H.SW_G224
    _write_case(str(tmp_path), band, atm, kd, block_size=8)
    r = subprocess.run([EXE, band, str(tmp_path)] + [os.path.join(H.NN_DIR, f) for f in files], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "mean of flux_down is:" in r.stdout
    ncol, nlay = atm["play"].shape
    up = np.fromfile(tmp_path / "flux_up.f32", np.float32).reshape(ncol, nlay + 1)
    dn = np.fromfile(tmp_path / "flux_dn.f32", np.float32).reshape(ncol, nlay + 1)
    k_dist = api.ty_gas_optics_rrtmgp(gpu_ctx); assert k_dist.load(kd) == ""
    dnets, onets = H.device_nets(gpu_ctx, files), H.oracle_nets(files)
    gc = H.gas_concs(atm["gases"])
    if band == "lw":
        pu, pd = api.lw_fluxes_host(k_dist, dnets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["sfc_emis"], gc, tlev=atm["tlev"],
                                    top_at_1=atm["top_at_1"])
        ref = O.gas_optics_lw(kd, onets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"])
        ru, rd = O.rte_lw(kd, atm["top_at_1"], ref["tau"], ref["lay_source"], ref["lev_source"], ref["sfc_source"],
                          np.repeat(atm["sfc_emis"][:, None], 16, 1))
        assert np.abs(up - ru).max() <= H.FLUX_TOL and np.abs(dn - rd).max() <= H.FLUX_TOL
    else:
        mu0 = np.where(atm["usecol"], atm["mu0"], -1.0).astype(np.float32)     # the host-buffer entry point marks night columns by mu0 <= 0
        pu, pd, _ = api.sw_fluxes_host(k_dist, dnets, atm["play"], atm["plev"], atm["tlay"], mu0, atm["sfc_alb"], gc, tsi=atm["tsi"],
                                       top_at_1=atm["top_at_1"])
        assert np.all(up[night] == 0) and np.all(dn[night] == 0) and up[day].max() > 0
    # the compiled driver (stage API, blocks of 8) and the Python mirror (fused whole-path entry point) run the same kernels
    assert np.abs(up - pu).max() <= 1e-5 * np.abs(pu).max() and np.abs(dn - pd).max() <= 1e-5 * np.abs(pd).max()
