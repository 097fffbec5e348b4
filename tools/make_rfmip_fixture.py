"""Build tests/golden/rfmip_inputs.npz from the reference's RFMIP input file (run in the build container only;
/root/reference does not exist on the GPU box).  Follows the reading/conditioning of the reference drivers:
  examples/rfmip-clear-sky/mo_rfmip_io.F90 (read_and_block_pt :185-260, read_and_block_gases_ty :451-680,
  read_and_block_lw_bc / _sw_bc), rrtmgp_rfmip_lw.F90:287,300-305, rrtmgp_rfmip_sw.F90:285-287.
Column index = iexp*100 + isite (experiment-major), i.e. the reference's (nlay, ncol, nexp) -> blocks reshape.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from nc4min import NC4File

SRC = "/root/reference/examples/rfmip-clear-sky/multiple_input4MIPs_radiation_RFMIP_UColorado-RFMIP-1-2_none.nc"
# NN input name -> RFMIP file variable (determine_gas_names, mo_rfmip_io.F90:317-411)
GAS_FILE_NAME = dict(co2="carbon_dioxide", n2o="nitrous_oxide", ch4="methane", co="carbon_monoxide", ccl4="carbon_tetrachloride",
                     cfc22="hcfc22", cfc11="cfc11", cfc12="cfc12", hfc143a="hfc143a", hfc125="hfc125", hfc23="hfc23",
                     hfc32="hfc32", hfc134a="hfc134a", cf4="cf4")

f = NC4File(SRC)
nexp, nsite, nlay = f.read("temp_layer").shape
rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (nexp,) + a.shape).reshape((nexp * nsite,) + a.shape[1:]))
flat = lambda a: np.ascontiguousarray(a.reshape((nexp * nsite,) + a.shape[2:]))
out = dict(
    p_lay=rep(f.read("pres_layer")), p_lev=rep(f.read("pres_level")),
    t_lay=flat(f.read("temp_layer")), t_lev=flat(f.read("temp_level")),
    h2o=flat(f.read("water_vapor")) * np.float32(float(f.attr_str("water_vapor", "units"))),
    o3=flat(f.read("ozone")) * np.float32(float(f.attr_str("ozone", "units"))),
    sfc_t=flat(f.read("surface_temperature")), sfc_emis=rep(f.read("surface_emissivity")),
    sfc_alb=rep(f.read("surface_albedo")), sza=rep(f.read("solar_zenith_angle")), tsi=rep(f.read("total_solar_irradiance")),
)
for nn_name, file_name in GAS_FILE_NAME.items():
    v = f.read(file_name + "_GM").astype(np.float32) * np.float32(float(f.attr_str(file_name + "_GM", "units")))
    out["gm_" + nn_name] = np.repeat(v, nsite).astype(np.float32)   # per column (constant within an experiment)
out = {k: v.astype(np.float32) for k, v in out.items()}
dst = os.path.join(ROOT, "tests", "golden", "rfmip_inputs.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst) / 1e6, "MB", {k: v.shape for k, v in out.items() if v.ndim > 1})
