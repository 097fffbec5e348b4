"""world_size-2 gloo test of the N>1 host logic: contiguous column shards + one all_gather of the fluxes reproduce
the single-process result exactly (per-rank fluxes come from the oracle here; on GPUs they come from the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers as H


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ncol, q):
    sys.path[:0] = [H.ROOT, os.path.join(H.ROOT, "oracle"), os.path.join(H.ROOT, "tests")]
    import torch
    import torch.distributed as dist
    import oracle as O
    from rte_rrtmgp_nn_b200 import sharding, spectral, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    atm = synth.make_atmosphere(ncol, 20, seed=3)
    c0, c1 = sharding.shard_bounds(ncol, rank, world)
    kd = spectral.synthetic_kdist_lw(128)
    nets = H.oracle_nets(H.LW_G128)
    g = {k: (v[c0:c1] if np.ndim(v) == 2 else v) for k, v in atm["gases"].items()}
    go = O.gas_optics_lw(kd, nets, atm["play"][c0:c1], atm["plev"][c0:c1], atm["tlay"][c0:c1], atm["tsfc"][c0:c1], g, tlev=atm["tlev"][c0:c1])
    emis = np.repeat(atm["sfc_emis"][c0:c1, None], 16, 1)
    up, dn = O.rte_lw(kd, True, go["tau"], go["lay_source"], go["lev_source"], go["sfc_source"], emis)
    full = sharding.gather_fluxes([torch.from_numpy(up), torch.from_numpy(dn)], ncol)
    if rank == 0:
        q.put((full[0].numpy(), full[1].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from rte_rrtmgp_nn_b200 import sharding
    for ncol in (1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            b = [sharding.shard_bounds(ncol, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == ncol
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [c1 - c0 for c0, c1 in b]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.max_shard(ncol, world)


@pytest.mark.timeout(180)
def test_two_rank_gloo_gather_matches_single_process():
    import torch.multiprocessing as mp
    import oracle as O
    from rte_rrtmgp_nn_b200 import spectral, synth
    ncol, world = 11, 2   # odd: shards of 5 and 6 columns
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncol, q)) for r in range(world)]
    for p in procs:
        p.start()
    up, dn = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    atm = synth.make_atmosphere(ncol, 20, seed=3)
    kd = spectral.synthetic_kdist_lw(128)
    nets = H.oracle_nets(H.LW_G128)
    go = O.gas_optics_lw(kd, nets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"])
    rup, rdn = O.rte_lw(kd, True, go["tau"], go["lay_source"], go["lev_source"], go["sfc_source"], np.repeat(atm["sfc_emis"][:, None], 16, 1))
    assert np.array_equal(up, rup) and np.array_equal(dn, rdn)
