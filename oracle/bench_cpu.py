"""ctypes front end of bench_cpu.c: the compiled CPU baseline (OpenMP over column blocks, blocked SGEMM over a block's
samples, gas-optics / solver split, best of N).  BENCHMARK INFRASTRUCTURE ONLY: bench.py's cpu_baseline and --impl reference
legs (and their CPU test) load it; the product never does."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class _Net(C.Structure):
    _fields_ = [("nlayers", C.c_int), ("dims", c_int_p), ("wpack", c_float_p), ("bpack", c_float_p), ("act", c_int_p),
                ("ymean", c_float_p), ("ystd", c_float_p), ("xmin", c_float_p), ("xmax", c_float_p)]


class _Problem(C.Structure):
    _fields_ = [("ncol", C.c_int), ("nlay", C.c_int), ("top_at_1", C.c_int),
                ("ngpt_lw", C.c_int), ("nbnd_lw", C.c_int), ("ntemp", C.c_int), ("band_lims_lw", c_int_p), ("totplnk", c_float_p),
                ("temp_ref_min", C.c_float), ("totplnk_delta", C.c_float), ("lw_tau", C.POINTER(_Net)), ("lw_pfrac", C.POINTER(_Net)),
                ("gas_lw", C.POINTER(c_float_p)), ("gas_mode_lw", c_int_p),
                ("ngpt_sw", C.c_int), ("solar_source", c_float_p), ("sw_abs", C.POINTER(_Net)), ("sw_ray", C.POINTER(_Net)),
                ("gas_sw", C.POINTER(c_float_p)), ("gas_mode_sw", c_int_p),
                ("play", c_float_p), ("plev", c_float_p), ("tlay", c_float_p), ("tlev", c_float_p), ("tsfc", c_float_p),
                ("sfc_emis", c_float_p), ("sfc_alb", c_float_p), ("mu0", c_float_p),
                ("lw_up", c_float_p), ("lw_dn", c_float_p), ("sw_up", c_float_p), ("sw_dn", c_float_p), ("sw_dir", c_float_p)]


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "_build", "libbench_cpu.so")
        srcs = [os.path.join(_HERE, f) for f in ("bench_cpu.c", "oracle.c")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(so)
        _lib.orcb_run.restype = C.c_int
        _lib.orcb_run.argtypes = [C.POINTER(_Problem), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_float_p)


class Problem:
    """One synthetic workload held in host memory: atmosphere + the four networks + the spectral tables."""

    def __init__(self, kd_lw, kd_sw, nets_lw, nets_sw, atm, want_fluxes=False):
        self.keep = []
        f32 = lambda a: self._k(np.ascontiguousarray(a, np.float32))
        ncol, nlay = atm["play"].shape
        P = _Problem()
        P.ncol, P.nlay, P.top_at_1 = ncol, nlay, int(bool(atm.get("top_at_1", True)))
        P.ngpt_lw, P.nbnd_lw, P.ntemp = int(kd_lw["ngpt"]), int(kd_lw["nbnd"]), int(np.shape(kd_lw["totplnk"])[-1])
        P.band_lims_lw = self._k(np.ascontiguousarray(kd_lw["band_lims_gpt"], np.int32)).ctypes.data_as(c_int_p)
        P.totplnk = _fp(f32(kd_lw["totplnk"]))
        P.temp_ref_min, P.totplnk_delta = float(kd_lw["temp_ref_min"]), float(kd_lw["totplnk_delta"])
        P.ngpt_sw = int(kd_sw["ngpt"])
        P.solar_source = _fp(f32(kd_sw["solar_source"]))
        self.nets = [self._net(n) for n in (nets_lw[0], nets_lw[1], nets_sw[0], nets_sw[1])]
        P.lw_tau, P.lw_pfrac, P.sw_abs, P.sw_ray = [C.pointer(n) for n in self.nets]
        P.gas_lw, P.gas_mode_lw = self._gases(nets_lw[0], atm["gases"], ncol, nlay)
        P.gas_sw, P.gas_mode_sw = self._gases(nets_sw[0], atm["gases"], ncol, nlay)
        for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0"):
            setattr(P, k, _fp(f32(atm[k])))
        self.fluxes = None
        if want_fluxes:
            self.fluxes = {k: np.zeros((ncol, nlay + 1), np.float32) for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir")}
            for k, v in self.fluxes.items():
                setattr(P, k, _fp(v))
        self.P = P
        self.ncol, self.nlay = ncol, nlay

    def _k(self, a):
        self.keep.append(a)
        return a

    def _net(self, n):
        s = _Net()
        s.nlayers = n.nlayers
        s.dims = self._k(np.ascontiguousarray(n.dims_a, np.int32)).ctypes.data_as(c_int_p)
        s.wpack, s.bpack = _fp(self._k(np.ascontiguousarray(n.wpack, np.float32))), _fp(self._k(np.ascontiguousarray(n.bpack, np.float32)))
        s.act = self._k(np.ascontiguousarray(n.act, np.int32)).ctypes.data_as(c_int_p)
        s.ymean = _fp(None if n.ymean is None else self._k(np.ascontiguousarray(n.ymean, np.float32)))
        s.ystd = _fp(None if n.ystd is None else self._k(np.ascontiguousarray(n.ystd, np.float32)))
        s.xmin, s.xmax = _fp(self._k(np.ascontiguousarray(n.xmin, np.float32))), _fp(self._k(np.ascontiguousarray(n.xmax, np.float32)))
        return self._k(s)

    def _gases(self, net, gases, ncol, nlay):
        """By name, as compute_nn_inputs (rrtmgp/mo_gas_optics_rrtmgp.F90:708-760): mode 0 scalar, 1 profile, 2 field, -1 absent."""
        nx = net.dims[0]
        ptrs = (c_float_p * 32)()
        modes = np.full(32, -1, np.int32)
        for i, name in enumerate(net.input_names):
            if i < 2 or name not in gases:
                continue
            v = np.asarray(gases[name], np.float32)
            if i < 4:
                v = np.ascontiguousarray(np.broadcast_to(v, (ncol, nlay)))
            v = self._k(np.ascontiguousarray(v).reshape(-1) if v.ndim == 0 else np.ascontiguousarray(v))
            v = self._k(np.atleast_1d(v))
            ptrs[i] = _fp(v)
            modes[i] = 2 if v.ndim == 2 else (1 if v.size == nlay and v.ndim == 1 and v.size > 1 else 0)
        self._k(ptrs); self._k(modes)
        return C.cast(ptrs, C.POINTER(c_float_p)), modes.ctypes.data_as(c_int_p)

    def run(self, block, repeats=5, lw=True, sw=True):
        """-> dict(seconds, columns_per_s, gas_optics_s, solver_s, threads) for the best of `repeats` passes."""
        out = (C.c_double * 4)()
        rc = lib().orcb_run(C.byref(self.P), int(block), int(repeats), int(lw), int(sw), out)
        if rc:
            raise RuntimeError("orcb_run failed")
        return {"block": int(block), "seconds": out[0], "columns_per_s": self.ncol / out[0], "gas_optics_s_per_thread": out[1],
                "solver_s_per_thread": out[2], "threads": int(out[3])}
