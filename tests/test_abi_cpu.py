"""CPU tests of the C-ABI library and the host logic: it loads, exports every declared symbol, reads the model files,
and fails loudly without a GPU (no CPU fallback)."""
import glob
import os
import re

import numpy as np
import pytest

import helpers as H
from rte_rrtmgp_nn_b200 import _lib, api


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = open(os.path.join(H.ROOT, "include", "rrnn.h")).read()
    names = re.findall(r"RRNN_API\s+[\w\s\*]+?\b(rrnn_\w+)\s*\(", hdr)
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rrnn.h but not exported"
        assert n in _lib._SIGS, f"{n} has no ctypes signature"
    assert L.rrnn_version() >= 100


def test_every_entry_point_cites_the_reference():
    hdr = open(os.path.join(H.ROOT, "include", "rrnn.h")).read()
    assert len(re.findall(r"\.F90:\d+", hdr)) >= 30


def test_product_never_touches_the_oracle():
    """The product (package sources, csrc) must not import, link or execute anything under oracle/."""
    pkg = os.path.join(H.ROOT, "rte_rrtmgp_nn_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
            src = open(path, errors="replace").read()
            for line in src.splitlines():
                code = line.split("#")[0] if path.endswith(".py") else line.split("//")[0]
                assert not re.search(r"\bimport\s+oracle\b|\bfrom\s+oracle\b|liboracle|oracle/_build", code), (path, line)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RRNNError, match="no CPU fallback"):
        api.Context(0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(H.NN_DIR, "*.nc"))))
def test_cxx_netcdf4_reader_matches_python_reader(path):
    """rrnn_model_load_netcdf (C++, product) against oracle/nc4min.py (independent numpy reader) on every shipped file."""
    import nc4min
    ref = nc4min.load_nn_model(path)
    m = api.rrtmgp_network_type(None).load_netcdf(path)
    assert m.dims == ref["dims"] and m.input_names == ref["input_names"] and m.activations == ref["activations"]
    for l in range(len(ref["W"])):
        assert np.array_equal(m.weights(l), ref["W"][l]) and np.array_equal(m.bias(l), ref["b"][l])
    assert np.array_equal(m.coeffs_input_min, ref["xmin"]) and np.array_equal(m.coeffs_input_max, ref["xmax"])
    if ref["ymean"] is None:
        assert m.coeffs_output_mean is None
    else:
        assert np.array_equal(m.coeffs_output_mean, ref["ymean"]) and np.array_equal(m.coeffs_output_std, ref["ystd"])


def test_shipped_model_inventory():
    """all models are softsign-softsign-linear; LW takes 18 inputs, SW 7; input order differs between
    g256 and g128 (co2 n2o ch4 vs co2 ch4 n2o) -> inputs must be mapped by name."""
    a = api.rrtmgp_network_type(None).load_netcdf(os.path.join(H.NN_DIR, H.LW_G256[0]))
    b = api.rrtmgp_network_type(None).load_netcdf(os.path.join(H.NN_DIR, H.LW_G128[0]))
    s = api.rrtmgp_network_type(None).load_netcdf(os.path.join(H.NN_DIR, H.SW_G224[0]))
    assert a.dims == [18, 58, 58, 256] and s.dims == [7, 16, 16, 224]
    assert a.input_names[:7] == ["tlay", "play", "h2o", "o3", "co2", "n2o", "ch4"]
    assert b.input_names[:7] == ["tlay", "play", "h2o", "o3", "co2", "ch4", "n2o"]
    assert a.activations == ["softsign", "softsign", "linear"]


def test_ascii_model_round_trip(tmp_path):
    """network_type%load ASCII format (neural/mod_network.F90:163-209) + scaling sidecar: save -> load is lossless."""
    src = api.rrtmgp_network_type(None).load_netcdf(os.path.join(H.NN_DIR, H.SW_G224[1]))
    mt, st = str(tmp_path / "model.txt"), str(tmp_path / "scaling.txt")
    src.save(mt, st)
    txt = open(mt).read().split()
    assert txt[0] == "4" and txt[1:5] == ["7", "16", "16", "224"] and txt[-3:] == ["softsign", "softsign", "linear"]
    back = api.rrtmgp_network_type(None).load(mt, st)
    assert back.dims == src.dims and back.input_names == src.input_names and back.activations == src.activations
    for l in range(3):
        assert np.array_equal(back.weights(l), src.weights(l)) and np.array_equal(back.bias(l), src.bias(l))
    assert np.array_equal(back.coeffs_output_std, src.coeffs_output_std)


def test_reader_error_messages(tmp_path):
    bad = tmp_path / "x.nc"
    bad.write_bytes(b"CDF\x01" + b"\0" * 100)
    with pytest.raises(_lib.RRNNError, match="not a netCDF-4"):
        api.rrtmgp_network_type(None).load_netcdf(str(bad))
    with pytest.raises(_lib.RRNNError, match="can't find file"):
        api.rrtmgp_network_type(None).load_netcdf(str(tmp_path / "missing.nc"))


def test_gas_concs_and_spectral_tables():
    gc = api.ty_gas_concs(["h2o", "co2"])
    assert gc.set_vmr("co2", 4e-4) == "" and "should be" in gc.set_vmr("co2", 1.5)
    from rte_rrtmgp_nn_b200 import spectral
    kd = spectral.synthetic_kdist_lw(256)
    assert kd["totplnk"].shape == (16, 196) and kd["band_lims_gpt"][-1, 1] == 256
    # sum over bands of the band-integrated Planck radiance ~ sigma T^4 / pi (bands cover 10-3250 cm-1)
    T = 288.0
    assert abs(np.pi * kd["totplnk"][:, 128].sum() / (5.670374e-8 * T ** 4) - 1) < 2e-3
    ks = spectral.synthetic_kdist_sw(224)
    assert abs(ks["solar_source"].sum() - 1361.0) < 1e-2


def test_solar_var_ind_interp_host_routine():
    """ty_solar_var (extensions/solar_variability/mo_solar_variability.F90:20-183), a HOST-only routine in the reference and here:
    the library / Python mirror against the oracle's step-by-step restatement (bit for bit) and against an independent formulation
    (piecewise-linear through the end points and the month centres) on the reference's own table."""
    import oracle as O
    from rte_rrtmgp_nn_b200 import api
    tab = api.load_solar_var_file(os.path.join(H.ROOT, "data", "solar_variability", "rrtmgp-solar-var-tables.nc"))
    assert tab.shape == (134, 2)
    sv = api.ty_solar_var()
    assert sv.solar_var_ind_interp(0.3) == ("", None, None)          # no table loaded: nothing computed, no message (as the reference)
    assert sv.load(tab) == ""
    n = tab.shape[0]
    xs = np.concatenate([[0.0, 1.0], 0.5 / (n - 2) * np.array([0.5, 1.0, 1.5]), 1 - 0.5 / (n - 2) * np.array([0.5, 1.0, 1.5]),
                         np.random.default_rng(3).uniform(0, 1, 200)]).astype(np.float32)
    knots = np.concatenate([[0.0], (np.arange(n - 2) + 0.5) / (n - 2), [1.0]])
    for x in xs:
        err, mg, sb = sv.solar_var_ind_interp(float(x))
        oerr, omg, osb = O.solar_var_ind_interp(tab, x)
        assert err == oerr == ""
        assert np.float32(mg) == omg and np.float32(sb) == osb, (x, mg, omg, sb, osb)
        assert abs(mg - np.interp(float(x), knots, tab[:, 0].astype(np.float64))) <= 2e-6 * abs(tab[:, 0]).max()
        assert abs(sb - np.interp(float(x), knots, tab[:, 1].astype(np.float64))) <= 2e-6 * abs(tab[:, 1]).max() + 1e-9
    for bad in (-0.01, 1.01):
        assert sv.solar_var_ind_interp(bad)[0] == "solar_var_ind_interp: solcycfrac out of range" == O.solar_var_ind_interp(tab, bad)[0]
    sv.finalize()
    assert sv.avgcyc_ind is None
