"""netCDF-4 files through the library's own reader / writer (rrnn_nc_*, csrc/nc4_io.cpp): the Python face of what the reference's
drivers do with examples/mo_simple_netcdf.F90 (read_field :34-96, var_exists :308-318, create_dim :320-343, create_var :345-380,
write_field :167-237).  No libnetcdf / libhdf5 / h5py involved; nothing here needs a GPU."""
import ctypes as C

import numpy as np

from . import _lib

vp = C.c_void_p


class NcFile:
    """A netCDF-4 file open for reading (mode "r") or being created (mode "w": written on close)."""

    def __init__(self, path, mode="r"):
        self.h = vp()
        self.mode = mode
        fn = _lib.lib().rrnn_nc_open if mode == "r" else _lib.lib().rrnn_nc_create
        _lib.check(fn(str(path).encode(), C.byref(self.h)))
        self._dims = {}

    # ---- reading
    def var_exists(self, name):
        return bool(_lib.lib().rrnn_nc_var_exists(self.h, name.encode()))

    def shape(self, name):
        nd = C.c_int(0)
        shp = (C.c_longlong * 8)()
        _lib.check(_lib.lib().rrnn_nc_inq_var(self.h, name.encode(), C.byref(nd), shp))
        return tuple(int(shp[k]) for k in range(nd.value))

    def read_field(self, name):
        """Any numeric variable as float32, in file order (the transpose of the Fortran array read_field returns)."""
        shp = self.shape(name)
        out = np.empty(shp, np.float32)
        _lib.check(_lib.lib().rrnn_nc_get_var_float(self.h, name.encode(), out.ctypes.data_as(_lib.c_float_p), out.size))
        return out

    def get_att(self, var, att):
        buf = C.create_string_buffer(512)
        _lib.check(_lib.lib().rrnn_nc_get_att_text(self.h, var.encode(), att.encode(), buf, 512))
        return buf.value.decode()

    # ---- writing
    def create_dim(self, name, length):
        i = C.c_int(-1)
        _lib.check(_lib.lib().rrnn_nc_def_dim(self.h, name.encode(), int(length), C.byref(i)))
        self._dims[name] = i.value
        return i.value

    def write_field(self, name, dim_names, values, units=None):
        """create_var + write_field: values in file order, shape = the lengths of dim_names."""
        a = np.ascontiguousarray(values, np.float32)
        ids = (C.c_int * len(dim_names))(*[self._dims[d] for d in dim_names])
        _lib.check(_lib.lib().rrnn_nc_put_var_float(self.h, name.encode(), len(dim_names), ids, a.ctypes.data_as(_lib.c_float_p),
                                                    None if units is None else units.encode()))

    def close(self):
        if self.h:
            h, self.h = self.h, vp()
            _lib.check(_lib.lib().rrnn_nc_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            if self.h and self.mode == "r":
                self.close()
        except Exception:
            pass
