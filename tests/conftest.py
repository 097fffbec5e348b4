import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rte_rrtmgp_nn_b200 import api
    return api.default_context(0)


@pytest.fixture(params=[1, 0], ids=["tensor_cores", "fp32_ffma"])
def nn_variant(request, gpu_ctx):
    """Both MLP variants of the NN gas optics: tcgen05 tensor cores (the default) and the fp32 FFMA kernel, each with
    the tau tolerance stated for it (helpers.TAU_RTOL_*).  The fixture checks WHICH kernel served the test's gas-optics
    calls (rrnn_ctx_nn_kernel_counts): a "tensor_cores" id fails if the FFMA kernel ran, and the other way round."""
    import helpers as H
    gpu_ctx.set_flag("nn_tensor_cores", request.param)
    tc0, ff0 = gpu_ctx.nn_kernel_counts
    yield (H.TAU_RTOL_TC if request.param else H.TAU_RTOL_FP32)
    tc1, ff1 = gpu_ctx.nn_kernel_counts
    gpu_ctx.set_flag("nn_tensor_cores", 1)
    if request.param:
        assert ff1 == ff0, f"a tensor_cores test ran the fp32 FFMA kernel {ff1 - ff0} time(s) (silent fallback)"
        assert tc1 > tc0, "a tensor_cores test never launched the tcgen05 kernel"
    else:
        assert tc1 == tc0, f"an fp32_ffma test ran the tcgen05 kernel {tc1 - tc0} time(s)"


@pytest.fixture(params=[(0, 1), (0, 0), (1, 0)], ids=["v7_wide_lw", "v6_tma_packed", "v3_scalar"])
def solver_variant(request, gpu_ctx):
    """The RTE solver kernels behind rrnn_lw_solver_noscat / rrnn_sw_solver_2stream: TMA-staged packed fp32x2 (rte_solvers_tma.cu;
    the default, with four g-points per lane in the LW solver where the shape fits -- lw_solver_v7 -- and with two -- lw_solver_v6)
    and one g-point per lane (rte_solvers.cu, the fallback for shapes the packed kernels do not take)."""
    gpu_ctx.set_flag("solver_variant", request.param[0])
    gpu_ctx.set_flag("solver_wide", request.param[1])
    yield request.param[0]
    gpu_ctx.set_flag("solver_variant", 0)
    gpu_ctx.set_flag("solver_wide", 1)
