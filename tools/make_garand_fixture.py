"""Build tests/golden/garand_atmos.npz from the reference's all-sky atmosphere (run in the build container only;
/root/reference does not exist on the GPU box).  Follows read_atmos, examples/all-sky/mo_garand_atmos_io.F90:41-88:
fields (lay|lev, col) transposed to (col, lay|lev); gases h2o co2 o3 n2o co ch4 o2 n2; optional col_dry."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from nc4min import NC4File

f = NC4File("/root/reference/examples/all-sky/garand-atmos-1.nc")
out = {k: np.ascontiguousarray(f.read(k).T.astype(np.float32)) for k in ("p_lay", "t_lay", "p_lev", "t_lev", "col_dry")}
for g in ("h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "n2"):
    out["vmr_" + g] = np.ascontiguousarray(f.read("vmr_" + g).T.astype(np.float32))
dst = os.path.join(ROOT, "tests", "golden", "garand_atmos.npz")
np.savez_compressed(dst, **out)
print(dst, {k: v.shape for k, v in out.items()})
