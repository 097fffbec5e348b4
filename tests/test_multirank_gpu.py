"""N ranks == 1 rank, bit for bit, on real GPUs: every rank runs the CUDA path on its
contiguous column shard, the fluxes are gathered with NCCL (rte_rrtmgp_nn_b200.sharding.gather_fluxes), and rank 0 compares the
gathered arrays with the same call over ALL columns on its own GPU.  Skipped on a one-GPU box (the driver's round-end test
box); run with `gpurun --gpus 2 -- python -m pytest tests/test_multirank_gpu.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _fluxes(ctx_dev, atm, c0, c1):
    """LW + SW fluxes of columns [c0, c1) through the fused device-buffer entry points; returns five CPU tensors."""
    import torch
    from rte_rrtmgp_nn_b200 import api, spectral
    dev = torch.device("cuda", ctx_dev)
    ctx = api.Context(ctx_dev)
    k_lw = api.ty_gas_optics_rrtmgp(ctx); assert k_lw.load(spectral.synthetic_kdist_lw(256)) == ""
    k_sw = api.ty_gas_optics_rrtmgp(ctx); assert k_sw.load(spectral.synthetic_kdist_sw(224)) == ""
    nl = H.device_nets(ctx, H.LW_G256); ns = H.device_nets(ctx, H.SW_G224)
    d = {k: torch.from_numpy(np.ascontiguousarray(atm[k][c0:c1])).to(dev) for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0")}
    gc = api.ty_gas_concs()
    for k, v in atm["gases"].items():
        gc.set_vmr(k, torch.from_numpy(np.ascontiguousarray(v[c0:c1])).to(dev) if np.ndim(v) == 2 else float(v))
    n, nlev = c1 - c0, atm["play"].shape[1] + 1
    out = [torch.empty((n, nlev), device=dev) for _ in range(5)]
    api.lw_fluxes(k_lw, nl, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gc, out[0], out[1], tlev=d["tlev"])
    api.sw_fluxes(k_sw, ns, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gc, out[2], out[3], out[4])
    torch.cuda.synchronize(dev)
    return out


def _worker(rank, world, port, ncol, nlay, q):
    sys.path[:0] = [H.ROOT, os.path.join(H.ROOT, "oracle"), os.path.join(H.ROOT, "tests")]
    import torch
    import torch.distributed as dist
    from rte_rrtmgp_nn_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    atm = synth.make_atmosphere(ncol, nlay, seed=31)          # the same columns on every rank
    c0, c1 = sharding.shard_bounds(ncol, rank, world)
    mine = _fluxes(rank, atm, c0, c1)
    full = sharding.gather_fluxes(mine, ncol)                  # NCCL all_gather over NVLink
    if rank == 0:
        alone = _fluxes(0, atm, 0, ncol)                       # one rank, all columns
        q.put([bool(torch.equal(a, b)) for a, b in zip(full, alone)] + [float(full[0].sum())])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_n_ranks_equal_one_rank_bit_for_bit():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    ncol, nlay = 4099, 60      # prime: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncol, nlay, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=500)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res[:5]), f"gathered fluxes of {world} ranks differ from the single-rank result: {res}"
    assert res[5] > 0
