"""ctypes binding of librrnn_b200.so (the C ABI declared in include/rrnn.h).

The library is built in-tree by rte_rrtmgp_nn_b200/csrc/Makefile (see __graft_entry__.build).  There is no
fallback: if the shared object is missing, import of the compute API fails loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# RRNN_LIB_PATH: load another build of the same library (kernel-variant experiments); never a different implementation
LIB_PATH = os.environ.get("RRNN_LIB_PATH") or os.path.join(_HERE, "lib", "librrnn_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
vp = C.c_void_p


class rrnn_gas_t(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("conc", vp), ("value", C.c_float), ("ndims", C.c_int)]


class RRNNError(RuntimeError):
    pass


def build(verbose=False):
    """Compile librrnn_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RRNNError("building librrnn_b200.so failed")
    return LIB_PATH


_lib = None

_SIGS = {
    "rrnn_last_error": (C.c_char_p, []),
    "rrnn_version": (C.c_int, []),
    "rrnn_device_count": (C.c_int, []),
    "rrnn_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "rrnn_ctx_destroy": (C.c_int, [vp]),
    "rrnn_ctx_set_stream": (C.c_int, [vp, vp]),
    "rrnn_ctx_stream": (vp, [vp]),
    "rrnn_ctx_synchronize": (C.c_int, [vp]),
    "rrnn_ctx_set_flag": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "rrnn_ctx_launch_count": (C.c_longlong, [vp]),
    "rrnn_ctx_last_nn_kernel": (C.c_int, [vp]),
    "rrnn_ctx_nn_kernel_counts": (C.c_int, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "rrnn_dev_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "rrnn_dev_free": (C.c_int, [vp, vp]),
    "rrnn_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rrnn_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rrnn_ctx_profile": (C.c_int, [vp, C.c_int]),
    "rrnn_ctx_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), c_int_p]),
    "rrnn_ctx_set_chunk_columns": (C.c_int, [vp, C.c_int]),
    "rrnn_model_load_netcdf": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "rrnn_model_load_ascii": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.POINTER(vp)]),
    "rrnn_model_save_ascii": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "rrnn_model_create": (C.c_int, [vp, C.c_int, c_int_p, c_float_p, c_float_p, c_int_p, c_float_p, c_float_p, c_float_p,
                                    c_float_p, C.c_char_p, C.POINTER(vp)]),
    "rrnn_model_destroy": (C.c_int, [vp]),
    "rrnn_model_nlayers": (C.c_int, [vp]),
    "rrnn_model_dims": (C.c_int, [vp, c_int_p]),
    "rrnn_model_input_name": (C.c_int, [vp, C.c_int, C.c_char_p]),
    "rrnn_model_activation": (C.c_int, [vp, C.c_int]),
    "rrnn_model_get": (C.c_int, [vp, C.c_int, C.c_int, c_float_p, c_int_p]),
    "rrnn_kdist_create": (C.c_int, [vp, C.c_int, C.c_int, c_int_p, C.c_int, c_float_p, C.c_float, C.c_float, c_float_p,
                                    C.POINTER(vp)]),
    "rrnn_kdist_destroy": (C.c_int, [vp]),
    "rrnn_kdist_set_tsi": (C.c_int, [vp, C.c_float]),
    "rrnn_kdist_set_solar_tables": (C.c_int, [vp, c_float_p, c_float_p, c_float_p]),
    "rrnn_kdist_set_solar_variability": (C.c_int, [vp, C.c_float, C.c_float, C.c_int, C.c_float]),
    "rrnn_solar_var_ind_interp": (C.c_int, [c_float_p, C.c_int, C.c_float, c_float_p, c_float_p]),
    "rrnn_kdist_get_solar_source": (C.c_int, [vp, c_float_p]),
    "rrnn_kdist_set_optimal_angle_fit": (C.c_int, [vp, c_float_p]),
    "rrnn_compute_optimal_angles": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "rrnn_sum_byband": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "rrnn_net_byband": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "rrnn_net_flux": (C.c_int, [vp, C.c_size_t, vp, vp, vp]),
    "rrnn_get_col_dry": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "rrnn_interp_tlev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "rrnn_compute_nn_inputs": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.POINTER(rrnn_gas_t), C.c_int, vp]),
    "rrnn_output_sgemm_tau": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp]),
    "rrnn_output_sgemm_pfrac": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "rrnn_output_sgemm_lw": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "rrnn_planck_source_nn": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp]),
    "rrnn_gas_optics_lw": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp,
                                     C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp]),
    "rrnn_gas_optics_lw_compact": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp,
                                             C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_lw_solver_noscat_compact": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, vp, vp, vp, vp,
                                                vp, vp, vp, vp]),
    "rrnn_gas_optics_sw": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, vp, vp, vp, C.POINTER(rrnn_gas_t), C.c_int,
                                     vp, vp, vp, vp]),
    "rrnn_lw_solver_noscat": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, vp, vp, vp,
                                        vp, vp, vp, vp, vp]),
    "rrnn_lw_solver_noscat_ext": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p] + [vp] * 15),
    "rrnn_lw_solver_2stream": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 11),
    "rrnn_rte_lw_2stream": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int] + [vp] * 11),
    "rrnn_rte_lw_ext": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 15),
    "rrnn_rte_lw": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_sw_solver_2stream": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_sw_solver_2stream_ext": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 14),
    "rrnn_rte_sw": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_cloud_lut_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                        c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.POINTER(vp)]),
    "rrnn_cloud_pade_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [c_float_p] * 12 + [C.POINTER(vp)]),
    "rrnn_cloud_lut_destroy": (C.c_int, [vp]),
    "rrnn_cloud_optics": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_sampled_mask": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "rrnn_draw_samples": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_delta_scale_2str": (C.c_int, [vp, C.c_size_t, vp, vp, vp]),
    "rrnn_increment_1scl_bybnd": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "rrnn_increment_2str_bybnd": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    "rrnn_heating_rate": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "rrnn_calc_heating_rate": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "rrnn_lw_fluxes_host": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                      vp, C.POINTER(rrnn_gas_t), C.c_int, vp, vp]),
    "rrnn_sw_fluxes_host": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                      C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp]),
    "rrnn_rte_lw_byband": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_rte_sw_byband": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_rte_lw_clouds": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_rte_sw_clouds": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_lw_fluxes_allsky": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                        C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp]),
    "rrnn_sw_fluxes_allsky": (C.c_int, [vp, vp, C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                        C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_lw_fluxes_allsky_host": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                             C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp]),
    "rrnn_sw_fluxes_allsky_host": (C.c_int, [vp, vp, C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                             C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "rrnn_nc_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "rrnn_nc_create": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "rrnn_nc_close": (C.c_int, [vp]),
    "rrnn_nc_var_exists": (C.c_int, [vp, C.c_char_p]),
    "rrnn_nc_inq_var": (C.c_int, [vp, C.c_char_p, c_int_p, C.POINTER(C.c_longlong)]),
    "rrnn_nc_get_var_float": (C.c_int, [vp, C.c_char_p, c_float_p, C.c_size_t]),
    "rrnn_nc_get_att_text": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "rrnn_nc_def_dim": (C.c_int, [vp, C.c_char_p, C.c_longlong, c_int_p]),
    "rrnn_nc_put_var_float": (C.c_int, [vp, C.c_char_p, C.c_int, c_int_p, c_float_p, C.c_char_p]),
    "rrnn_multi_create": (C.c_int, [C.c_int, c_int_p, C.POINTER(vp)]),
    "rrnn_multi_destroy": (C.c_int, [vp]),
    "rrnn_multi_ndev": (C.c_int, [vp]),
    "rrnn_multi_ctx": (vp, [vp, C.c_int]),
    "rrnn_multi_set_flag": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "rrnn_multi_model_load_netcdf": (C.c_int, [vp, C.c_char_p, c_int_p]),
    "rrnn_multi_kdist_create": (C.c_int, [vp, C.c_int, C.c_int, c_int_p, C.c_int, c_float_p, C.c_float, C.c_float, c_float_p, c_int_p]),
    "rrnn_multi_kdist_set_tsi": (C.c_int, [vp, C.c_int, C.c_float]),
    "rrnn_multi_lw_fluxes_host": (C.c_int, [vp, C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                            C.POINTER(rrnn_gas_t), C.c_int, vp, vp]),
    "rrnn_multi_sw_fluxes_host": (C.c_int, [vp, C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                            C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp]),
    "rrnn_lw_fluxes": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                 C.POINTER(rrnn_gas_t), C.c_int, vp, vp]),
    "rrnn_sw_fluxes": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                 C.POINTER(rrnn_gas_t), C.c_int, vp, vp, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RRNNError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RRNNError(lib().rrnn_last_error().decode(errors="replace"))


def last_error():
    return lib().rrnn_last_error().decode(errors="replace")
