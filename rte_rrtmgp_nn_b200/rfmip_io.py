"""RFMIP clear-sky input / output files: examples/rfmip-clear-sky/mo_rfmip_io.F90 on top of the library's netCDF-4 reader / writer.

  read_size                 :49-67     (ncol, nlay, nexp) = (site, layer, expt)
  read_and_block_pt         :74-209    p_lay, p_lev (site-only in the file, replicated over experiments), t_lay, t_lev
  read_and_block_lw_bc      :250-307   surface emissivity, surface temperature
  read_and_block_sw_bc      :214-245   surface albedo, total solar irradiance, solar zenith angle
  determine_gas_names       :317-411   NN / k-distribution gas name -> variable name in the RFMIP file
  read_and_block_gases_ty   :451-680   h2o / o3 per (expt, site, layer) and the well-mixed gases per experiment, each times its
                                       `units` scaling factor (read_scaling :683-698)
  unblock_and_write         :734-870   (ncol*nexp, nlev) fluxes -> a variable (expt, site, level)
Arrays come back column-major over experiments -- column index = iexp * nsite + isite -- with layers last, i.e. the C view of the
reference's (nlay, blocksize, nblocks) blocks laid end to end; `blocksize` therefore is a property of the caller's loop
(drivers.rrtmgp_rfmip_lw), not of the arrays.  The input file ships under data/rfmip (a copy of the reference's, as data)."""
import os

import numpy as np

from .ncio import NcFile

_HERE = os.path.dirname(os.path.abspath(__file__))
RFMIP_FILE = os.path.join(os.path.dirname(_HERE), "data", "rfmip", "multiple_input4MIPs_radiation_RFMIP_UColorado-RFMIP-1-2_none.nc")

# determine_gas_names, forcing_index = 1 (all available greenhouse gases), mo_rfmip_io.F90:330-395: the names the NN models use
# (neural/data model files, input_names) -> names in the RFMIP concentration file; h2o and o3 are profiles, the rest "_GM" scalars
CHEM_NAME = ("co", "ch4", "o2", "n2o", "n2", "co2", "ccl4", "ch4", "ch3br", "ch3cl", "cfc22")
CONC_NAME = ("carbon_monoxide", "methane", "oxygen", "nitrous_oxide", "nitrogen", "carbon_dioxide", "carbon_tetrachloride", "methane",
             "methyl_bromide", "methyl_chloride", "hcfc22")
GM_GASES = ("co2", "n2o", "ch4", "co", "ccl4", "cfc22", "cfc11", "cfc12", "hfc143a", "hfc125", "hfc23", "hfc32", "hfc134a", "cf4")


def determine_gas_names(names_in_kdist=("h2o", "o3") + GM_GASES):
    """names_in_kdist -> names_in_file (mo_rfmip_io.F90:397-406: chemical names are replaced, everything else keeps its name)."""
    table = dict(zip(CHEM_NAME, CONC_NAME))
    table.update(h2o="water_vapor", o3="ozone")
    return [table.get(g, g) for g in names_in_kdist]


def read_size(file_name=RFMIP_FILE):
    with NcFile(file_name) as f:
        nexp, ncol, nlay = f.shape("temp_layer")
    return ncol, nlay, nexp


def _rep(a, nexp):    # a field without the experiment dimension: the same for every experiment (spread(..., ncopies = nexp_l))
    return np.ascontiguousarray(np.broadcast_to(a[None], (nexp,) + a.shape).reshape((nexp * a.shape[0],) + a.shape[1:]))


def _flat(a):         # (expt, site, ...) -> (expt*site, ...)
    return np.ascontiguousarray(a.reshape((a.shape[0] * a.shape[1],) + a.shape[2:]))


def read_and_block_pt(file_name=RFMIP_FILE):
    ncol, nlay, nexp = read_size(file_name)
    with NcFile(file_name) as f:
        return (_rep(f.read_field("pres_layer"), nexp), _rep(f.read_field("pres_level"), nexp),
                _flat(f.read_field("temp_layer")), _flat(f.read_field("temp_level")))


def read_and_block_lw_bc(file_name=RFMIP_FILE):
    ncol, nlay, nexp = read_size(file_name)
    with NcFile(file_name) as f:
        return _rep(f.read_field("surface_emissivity"), nexp), _flat(f.read_field("surface_temperature"))


def read_and_block_sw_bc(file_name=RFMIP_FILE):
    ncol, nlay, nexp = read_size(file_name)
    with NcFile(file_name) as f:
        return (_rep(f.read_field("surface_albedo"), nexp), _rep(f.read_field("total_solar_irradiance"), nexp),
                _rep(f.read_field("solar_zenith_angle"), nexp))


def read_scaling(f, var_name):
    """read_scaling (mo_rfmip_io.F90:683-698): the `units` attribute holds the factor that turns the stored values into mole fractions."""
    return np.float32(float(f.get_att(var_name, "units")))


def read_and_block_gases_ty(file_name=RFMIP_FILE, gas_names=("h2o", "o3") + GM_GASES):
    """-> {gas: (ncol*nexp, nlay) float32 volume mixing ratios}; well-mixed gases are constant within a column."""
    ncol, nlay, nexp = read_size(file_name)
    names_in_file = determine_gas_names(gas_names)
    out = {}
    with NcFile(file_name) as f:
        for g, fn in zip(gas_names, names_in_file):
            if g in ("h2o", "o3"):
                out[g] = _flat(f.read_field(fn)) * read_scaling(f, fn)
            else:
                v = f.read_field(fn + "_GM").astype(np.float32) * read_scaling(f, fn + "_GM")      # (expt)
                per_col = np.repeat(v, ncol).astype(np.float32)
                out[g] = np.ascontiguousarray(np.broadcast_to(per_col[:, None], (ncol * nexp, nlay))).astype(np.float32)
    return out


def unblock_and_write(file_name, var_names, values, nexp, nsite, units="W m-2"):
    """unblock_and_write_3D (:769-799): fluxes (nexp*nsite, nlev) -> variables (expt, site, level) of a NEW netCDF-4 file (the
    reference writes into the template files that ship with RFMIP; creating the file here keeps the run self-contained)."""
    nlev = np.asarray(values[0]).shape[1]
    with NcFile(file_name, "w") as f:
        f.create_dim("expt", nexp); f.create_dim("site", nsite); f.create_dim("level", nlev)
        for name, a in zip(var_names, values):
            f.write_field(name, ("expt", "site", "level"), np.asarray(a, np.float32).reshape(nexp, nsite, nlev), units=units)
