"""The Fortran side (fortran/*.F90) cannot be compiled in this image (no Fortran compiler), so it is held to the C header by parsing:

* fortran/mo_rrnn_c_binding.F90 is GENERATED from include/rrnn.h (tools/gen_fortran_binding.py): it must be current, must carry one
  interface block per exported entry point, and every block's bind(C) name, argument count and argument kinds must match the header;
* every C function called from the veneer / driver modules must exist in the binding, be called with the right number of arguments,
  and each actual argument must look like the kind the binding declares (a c_ptr expression where a pointer goes by value, an
  integer(c_int) expression where an int goes, ...);
* the public procedures keep the reference's argument lists (names and order), cited per procedure below.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_fortran_binding as G   # noqa: E402

FORTRAN = os.path.join(ROOT, "fortran")
CALLERS = ["mo_rrnn_veneer.F90", "mo_rrnn_drivers.F90"]


def logical_lines(path):
    """Fortran free-form source -> statements: comments stripped (outside strings), continuation lines joined."""
    out, cur = [], ""
    for raw in open(path):
        line, q = "", None
        for ch in raw.rstrip("\n"):
            if q:
                line += ch
                if ch == q:
                    q = None
            elif ch in "\"'":
                q = ch
                line += ch
            elif ch == "!":
                break
            else:
                line += ch
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        out.append(cur + line)
        cur = ""
    assert cur == "", f"{path}: dangling continuation"
    return out


def split_args(s):
    """top-level comma split of the text between a call's parentheses"""
    args, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "\"'":
            q = ch
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def find_calls(stmt):
    """[(rrnn_name, [actual arguments])] for every rrnn_*( ... ) in one statement (nested calls included)"""
    calls = []
    for m in re.finditer(r"\b(rrnn_\w+)\s*\(", stmt):
        i, depth = m.end(), 1
        while depth and i < len(stmt):
            depth += {"(": 1, ")": -1}.get(stmt[i], 0)
            i += 1
        assert depth == 0, stmt
        calls.append((m.group(1), split_args(stmt[m.end():i - 1])))
    return calls


def binding_blocks():
    """name -> [(arg name, declaration)] parsed back from the GENERATED Fortran file (not from the generator's tables)"""
    blocks, cur = {}, None
    for st in logical_lines(os.path.join(FORTRAN, "mo_rrnn_c_binding.F90")):
        m = re.match(r"function (rrnn_\w+)\s*\((.*?)\)\s*bind\(C, name=\"(\w+)\"\) result\(rc\)", st)
        if m:
            assert m.group(1) == m.group(3), st
            cur = {"args": [a.strip() for a in m.group(2).split(",") if a.strip()], "decl": {}}
            blocks[m.group(1)] = cur
            continue
        if cur is not None and "::" in st and not st.startswith("import"):
            decl, names = st.split("::")
            for n in names.split(","):
                cur["decl"][re.sub(r"\(.*\)", "", n).strip()] = (decl.strip(), "(*)" in n or "(32)" in n)
        if st.startswith("end function"):
            cur = None
    return blocks


def test_binding_is_generated_from_the_header_and_current():
    assert open(os.path.join(FORTRAN, "mo_rrnn_c_binding.F90")).read() == G.generate(), \
        "fortran/mo_rrnn_c_binding.F90 is stale: run python tools/gen_fortran_binding.py"


def test_every_export_has_an_interface_block_with_matching_arguments():
    decls = G.parse_header()
    blocks = binding_blocks()
    assert len(decls) >= 85 and set(blocks) - {"rrnn_error_msg"} == {d[0] for d in decls}
    for name, ret, args, _ in decls:
        b = blocks[name]
        assert len(b["args"]) == len(args), name
        for (ctype, cname), fname in zip(args, b["args"]):
            decl, is_array = b["decl"][fname]
            t = ctype.replace("const ", "").replace(" const", "").strip()
            if t in ("int", "float", "size_t", "long long"):
                kind = {"int": "integer(c_int)", "float": "real(c_float)", "size_t": "integer(c_size_t)", "long long": "integer(c_long_long)"}[t]
                assert decl == f"{kind}, value" and not is_array, (name, cname, decl)
            elif t == "char*":
                assert decl == "character(kind=c_char)" and is_array, (name, cname, decl)
            elif t in ("int*", "long long*", "double*"):
                assert "value" not in decl and is_array, (name, cname, decl)
            elif t == "rrnn_gas_t*":
                assert decl == "type(rrnn_gas_t)" and is_array, (name, cname, decl)
            elif t.endswith("**") or t.endswith("* *"):
                assert decl == "type(c_ptr)", (name, cname, decl)           # by reference (handle out / handle array)
            else:
                assert decl == "type(c_ptr), value", (name, cname, decl)
        assert "rc" in b["decl"]


def looks_like(actual, decl, is_array):
    a = actual.strip()
    if decl == "type(c_ptr), value":
        return bool(re.fullmatch(r"c_loc\(.+\)|c_null_ptr|rrnn_ctx\(\)|the_ctx|p_\w+|host|\w+(%\w+)*%(p|kd|h|handle|lut)|this%(p|kd|h|lut)", a))
    if decl == "integer(c_int), value":
        return bool(re.fullmatch(r"int\(.+, c_int\)|merge\(1_c_int, 0_c_int, .+\)|\d+_c_int|t1", a))
    if decl == "integer(c_size_t), value":
        return bool(re.fullmatch(r"nlev_all( \* ngpt)?|\w+(%\w+)*%n|n \* c_sizeof\(1\.0_wp\)|min\(n, this%n\) \* c_sizeof\(1\.0_wp\)", a))
    if decl == "real(c_float), value":
        return bool(re.fullmatch(r"[\w%]+|\d*\._wp|\d+\.\d*_wp", a)) and not a.startswith("c_loc")
    if decl == "character(kind=c_char)":
        return a.startswith("c_str(")
    if decl == "type(c_ptr)":                       # by reference: a c_ptr variable / component / array of handles
        return bool(re.fullmatch(r"models|the_ctx|\w+(%\w+)*%(p|kd|h|handle|lut)", a))
    if decl == "type(rrnn_gas_t)":
        return a == "gases"
    if decl.startswith("integer(c_int)") and is_array:
        return bool(re.fullmatch(r"lims|dev|ids|id", a))
    return True


def test_fortran_callers_match_the_binding():
    blocks = binding_blocks()
    helper = {"rrnn_error_msg": 1, "rrnn_ctx": 0, "rrnn_select_device": 1, "rrnn_shutdown": 0, "rrnn_fill_zero": 1}
    own = {"rrnn_lw", "rrnn_sw", "rrnn_lw_multi", "rrnn_sw_multi", "rrnn_lw_allsky", "rrnn_sw_allsky"}
    ncalls = 0
    for f in CALLERS:
        for st in logical_lines(os.path.join(FORTRAN, f)):
            if re.match(r"(public|use|function|subroutine|end|procedure|private)\b", st):
                continue
            for name, actual in find_calls(st):
                if name in helper:
                    assert len(actual) == helper[name], (f, st)
                    continue
                if name in own:
                    continue
                assert name in blocks, f"{f}: {name} is not an entry point of include/rrnn.h"
                b = blocks[name]
                assert len(actual) == len(b["args"]), f"{f}: {name} called with {len(actual)} arguments, the header has {len(b['args'])}: {st}"
                for a, fname in zip(actual, b["args"]):
                    decl, is_array = b["decl"][fname]
                    assert looks_like(a, decl, is_array), f"{f}: {name}({fname}) takes '{decl}', got '{a}'"
                ncalls += 1
    assert ncalls >= 40


# The reference's argument lists (names, order).  Sources: rrtmgp/mo_gas_optics_rrtmgp.F90:239-243 (gas_optics_int), :433-437
# (gas_optics_ext), rte/mo_rte_lw.F90:60-64, rte/mo_rte_sw.F90:48-52, extensions/mo_heating_rates.F90:26,
# neural/mod_network_rrtmgp.F90:58 (load_netcdf), extensions/cloud_optics/mo_cloud_optics.F90:354-357 (cloud_optics),
# rrtmgp/mo_gas_concentrations.F90:91, 130 (init, set_vmr), rrtmgp/mo_gas_optics_rrtmgp.F90:1097 (set_tsi).
REFERENCE_SIGNATURES = {
    "gas_optics_int": "this play plev tlay tsfc gas_desc optical_props sources col_dry tlev neural_nets",
    "gas_optics_ext": "this play plev tlay gas_desc optical_props toa_src col_dry neural_nets",
    "rte_lw": "optical_props top_at_1 sources sfc_emis fluxes inc_flux n_gauss_angles use_2stream lw_Ds flux_up_Jac flux_dn_Jac",
    "rte_sw": "atmos top_at_1 mu0 inc_flux sfc_alb_dir_gpt sfc_alb_dif_gpt fluxes inc_flux_dif",
    "compute_heating_rate": "flux_up flux_dn plev heating_rate",
    "load_netcdf": "self filename",
    "cloud_optics": "this clwp ciwp reliq reice optical_props",
    "init": "this gas_names",
    "set_vmr_scalar": "this gas w",
    "set_vmr_1d": "this gas w",
    "set_vmr_2d": "this gas w",
    "set_tsi": "this tsi",
    "set_solar_variability": "this mg_index sb_index tsi",
    "rte_rrtmgp_config_checks_each": "extents values",
    "rte_rrtmgp_config_checks_all": "do_checks",
    "solar_var_ind_interp": "this solcycfrac mg_index sb_index",
}
REFERENCE_FILES = {
    "gas_optics_int": "rrtmgp/mo_gas_optics_rrtmgp.F90", "gas_optics_ext": "rrtmgp/mo_gas_optics_rrtmgp.F90", "rte_lw": "rte/mo_rte_lw.F90",
    "rte_sw": "rte/mo_rte_sw.F90", "compute_heating_rate": "extensions/mo_heating_rates.F90", "load_netcdf": "neural/mod_network_rrtmgp.F90",
    "set_tsi": "rrtmgp/mo_gas_optics_rrtmgp.F90", "set_solar_variability": "rrtmgp/mo_gas_optics_rrtmgp.F90",
    "rte_rrtmgp_config_checks_each": "rte/mo_rte_rrtmgp_config.F90", "rte_rrtmgp_config_checks_all": "rte/mo_rte_rrtmgp_config.F90",
    "solar_var_ind_interp": "extensions/solar_variability/mo_solar_variability.F90",
}


def signatures(path):
    sig = {}
    for st in logical_lines(path):
        m = re.match(r"(?:pure\s+|elemental\s+)?(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)", st, re.I)
        if m:
            sig.setdefault(m.group(1), " ".join(a.strip() for a in m.group(2).split(",") if a.strip()))
    return sig


def test_public_procedures_keep_the_reference_argument_lists():
    sig = signatures(os.path.join(FORTRAN, "mo_rrnn_veneer.F90"))
    for name, want in REFERENCE_SIGNATURES.items():
        assert sig.get(name, "").lower() == want.lower(), (name, sig.get(name))
    ref = "/root/reference"
    if os.path.isdir(ref):      # the build container has the reference: hold the table above to it, too
        for name, rel in REFERENCE_FILES.items():
            rsig = signatures(os.path.join(ref, rel))
            assert rsig[name].lower() == REFERENCE_SIGNATURES[name].lower(), (name, rsig[name])


def test_veneer_declares_the_reference_modules_and_types():
    src = "\n".join(re.sub(r"[ \t]+", " ", st) for st in logical_lines(os.path.join(FORTRAN, "mo_rrnn_veneer.F90"))).lower()
    for mod in ("mo_rte_kind", "mod_network_rrtmgp", "mo_gas_concentrations", "mo_optical_props", "mo_source_functions", "mo_fluxes",
                "mo_gas_optics_rrtmgp", "mo_rte_lw", "mo_rte_sw", "mo_heating_rates", "mo_cloud_optics", "mo_solar_variability"):
        assert re.search(rf"^module {mod}$", src, re.M), mod
        assert re.search(rf"^end module {mod}$", src, re.M), mod
    for ty in ("rrtmgp_network_type", "ty_gas_concs", "ty_optical_props_1scl", "ty_optical_props_2str", "ty_source_func_lw",
               "ty_fluxes_broadband", "ty_fluxes_flexible", "ty_gas_optics_rrtmgp", "ty_cloud_optics", "ty_solar_var"):
        assert re.search(rf"^type(, [^:]*)? :: {ty}$", src, re.M), ty
    assert re.search(r"generic, public :: gas_optics => gas_optics_int, gas_optics_ext", src)
    # balanced program units
    for kw in ("function", "subroutine", "module", "type", "interface"):
        opens = len(re.findall(rf"^(?:pure |elemental |abstract |logical |pure integer )*{kw}\b(?! ::|\()", src, re.M))
        if kw == "module":
            opens = len(re.findall(r"^module (?!procedure)\w+$", src, re.M))
        if kw == "type":
            opens = len(re.findall(r"^type(, [^:]*)? :: \w+$", src, re.M))
        closes = len(re.findall(rf"^end {kw}\b", src, re.M))
        assert opens == closes, (kw, opens, closes)


@pytest.mark.skipif(not any(os.path.exists(os.path.join(p, c)) for p in os.environ.get("PATH", "").split(":") for c in ("gfortran", "nvfortran", "ifx")),
                    reason="no Fortran compiler in this image")
def test_veneer_compiles_where_a_compiler_exists(tmp_path):
    subprocess.check_call(["make", "-C", FORTRAN, "clean", "all"])
