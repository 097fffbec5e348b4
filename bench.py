#!/usr/bin/env python
"""bench.py -- LW+SW flux columns/sec (NN gas optics + RTE) on N B200s, next to the CPU restatement.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3]): GCM-scale synthetic clear-sky LW+SW, 1,000,000 columns x 137 layers, the
g256 (LW) / g224 (SW) networks of the reference, columns sharded contiguously over the ranks (strong scaling:
the total is fixed).  One step = one pass of the hot path over all columns of the rank:
    gas_optics(neural_nets=) -> rte_lw      and      gas_optics(neural_nets=) -> boundary conditions -> rte_sw
  value : whole-job columns/s with inputs resident in HBM (rrnn_lw_fluxes + rrnn_sw_fluxes on device pointers),
          timed with CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks;
          for N>1 the final NCCL all_gather of the broadband fluxes is inside the timed step.
  e2e   : the same pass through the host-buffer C-ABI calls (rrnn_{lw,sw}_fluxes_host): pinned host inputs,
          H2D and D2H copies inside the timed region.
  roofline     : the dominant kernel (largest share of device time), timed live with CUDA events inside the lib.
  cpu_baseline : the oracle (C restatement of the reference kernels, -O3/AVX2/OpenMP) on a bounded column sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

NCOL_TOTAL = 1_000_000
NLAY = 137
NGPT_LW, NGPT_SW = 256, 224
LW_FILES = ("lw-g256-2018-12-04_absorption_58_58.nc", "lw-g256-2018-12-04_planck_frac_16_16.nc")
SW_FILES = ("sw-g224-2018-12-04-absorption_16_16.nc", "sw-g224-2018-12-04-rayleigh_16_16.nc")
NN_DIR = os.path.join(ROOT, "data", "nn")
NUNIQUE = 8192  # distinct synthetic columns, tiled to the full size (values do not change the work done)


def make_inputs(ncol, nlay, seed=12345):
    """Synthetic profiles of the named shape (rte_rrtmgp_nn_b200.synth), tiled from NUNIQUE distinct columns."""
    from rte_rrtmgp_nn_b200 import synth
    base = synth.make_atmosphere(min(NUNIQUE, ncol), nlay, seed=seed)
    reps = -(-ncol // base["play"].shape[0])

    def tile(a):
        return np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:ncol])
    atm = {k: tile(v) for k, v in base.items() if isinstance(v, np.ndarray)}
    atm["gases"] = {k: (tile(v) if np.ndim(v) == 2 else v) for k, v in base["gases"].items()}
    atm["top_at_1"] = True
    return atm


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes_per_column(nlay, nx_lw=18, nx_sw=7, lw_compact=True):
    """bytes each kernel must move per column (fp32), every array touched once.
    lw_compact: the LW sources cross HBM factored -- tau, pfrac (G,L) and two band tables (16,L), (16,L+1) instead of
    tau, lay_source (G,L) and lev_source (G,L+1); the denominators shrink with the traffic (8 instead of 12 B per g-point
    and layer), so `frac` stays a statement about the bytes that really have to move."""
    L, G, H = nlay, NGPT_LW, NGPT_SW
    lw_arrays = 4 * G * (2 * L) + 4 * 16 * (2 * L + 1) if lw_compact else 4 * G * (2 * L + (L + 1))
    return {
        "gas_optics_lw": 4 * ((nx_lw + 1) * L + 2 * (L + 1) + 1) + lw_arrays + 4 * G * 2,
        "lw_solver": lw_arrays + 4 * G * 2 + 8 * (L + 1),
        # g is identically zero on the NN path and is neither written nor read (2 arrays instead of 3)
        "gas_optics_sw": 4 * (nx_sw + 1) * L + 4 * H * (2 * L),
        "sw_solver": 4 * H * (2 * L + 3) + 4 + 12 * (L + 1),
    }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, ncol_per_launch):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture (profiles/*_ncu_traffic.json),
    scaled from the captured launch size to this run's columns per launch."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        d = json.load(f)
    k = d["kernels"].get(kernel)
    if not k:
        return None, None
    scale = ncol_per_launch / k["ncol_per_launch"]
    return (k["dram_bytes_read"] + k["dram_bytes_write"]) * scale, os.path.basename(files[-1])


def cpu_baseline(ncol_sample, nlay, threads_note=True):
    """Oracle (-O3/AVX2/OpenMP build) on a bounded sample of the same workload -> columns/s."""
    import nc4min
    import oracle as O
    from rte_rrtmgp_nn_b200 import spectral
    atm = make_inputs(ncol_sample, nlay, seed=777)
    kd, ks = spectral.synthetic_kdist_lw(NGPT_LW), spectral.synthetic_kdist_sw(NGPT_SW)
    lw = [O.Net(nc4min.load_nn_model(os.path.join(NN_DIR, f))) for f in LW_FILES]
    sw = [O.Net(nc4min.load_nn_model(os.path.join(NN_DIR, f))) for f in SW_FILES]
    emis = np.repeat(atm["sfc_emis"][:, None], kd["nbnd"], 1)
    alb = np.repeat(atm["sfc_alb"][:, None], NGPT_SW, 1)

    def one_pass():
        # the reference drivers loop over column blocks (OpenMP over blocks); the oracle threads over columns
        blk = 512
        for c0 in range(0, ncol_sample, blk):
            sl = slice(c0, min(ncol_sample, c0 + blk))
            g = {k: (v[sl] if np.ndim(v) == 2 else v) for k, v in atm["gases"].items()}
            go = O.gas_optics_lw(kd, lw, atm["play"][sl], atm["plev"][sl], atm["tlay"][sl], atm["tsfc"][sl], g, tlev=atm["tlev"][sl], fast=True)
            O.rte_lw(kd, True, go["tau"], go["lay_source"], go["lev_source"], go["sfc_source"], emis[sl], fast=True)
            gs = O.gas_optics_sw(ks, sw, atm["play"][sl], atm["plev"][sl], atm["tlay"][sl], g, fast=True)
            O.rte_sw(True, atm["mu0"][sl], gs["toa_src"], alb[sl], alb[sl], gs["tau"], gs["ssa"], gs["g"], fast=True)
    one_pass()  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    one_pass()
    dt = time.perf_counter() - t0
    return ncol_sample / dt, O.num_threads(True), dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is Fortran and cannot be
    built in this image (no Fortran compiler), so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is entitled to every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    ncol_sample = args.cpu_columns
    vals = []
    for i in range(args.warmup + args.steps):
        v, threads, dt = cpu_baseline(ncol_sample, NLAY)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "LW+SW flux columns/sec (NN gas optics + RTE)", "value": value, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, NCOL_TOTAL),
        "cpu_baseline": {"value": value, "unit": "columns/s", "cores": threads, "kind": "port",
                         "sample": f"{ncol_sample} columns x {NLAY} layers of the same synthetic workload per step; oracle/oracle.c "
                                   "(C restatement of the reference kernels, -O3 AVX2 OpenMP) -- the Fortran reference cannot be built here"},
        "e2e": {"value": value, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(ngpus, ncol_total):
    return {"workload": f"GCM-scale synthetic clear-sky LW+SW: {ncol_total} columns x {NLAY} layers, NN gas optics g256 (LW 18-58-58-256 + "
                        f"18-16-16-256) / g224 (SW 7-16-16-224 x2) + rte_lw (1 angle) + rte_sw (two-stream)",
            "ncol_total": ncol_total, "nlay": NLAY, "ngpt_lw": NGPT_LW, "ngpt_sw": NGPT_SW,
            "sharding": f"columns split contiguously over {ngpus} rank(s)",
            "l2": "per-step working set (optical properties, ~1.6 MB/column) is far larger than the 126 MB L2; no explicit flush"}


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set in the environment): from here on file descriptor 1 goes to stderr and the JSON line is written to
    the real stdout kept aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    global NLAY
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=NCOL_TOTAL, help="total columns (all ranks)")
    ap.add_argument("--nlay", type=int, default=NLAY, help="layers (137 = the headline configuration; 91 = the size sweep, configs[4])")
    ap.add_argument("--cpu-columns", type=int, default=65536, help="columns in the bounded CPU-baseline sample (~12 s per pass on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--device-inputs", action="store_true",
                    help="tile the NUNIQUE distinct columns on the device instead of on the host (the 1e7-column point of the size "
                         "sweep: no 40 GB of pinned host memory); implies --no-e2e")
    ap.add_argument("--fast-math", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--solver-buffer", type=int, default=0, help="0 auto, 1 shared memory, 2 L2-resident global scratch")
    ap.add_argument("--sw-fast-math", type=int, default=1, help="library default: the SW solver without Newton refinements (no measurable accuracy cost, DESIGN.md section 3b)")
    ap.add_argument("--solver-variant", type=int, default=0, help="0 packed two-g-points-per-lane solvers, 1 one g-point per lane")
    ap.add_argument("--solver-scratch-mb", type=int, default=0, help="L2 budget of the packed solvers' reverse-sweep scratch (0 = default)")
    ap.add_argument("--lw-compact-source", type=int, default=1, help="1 = LW sources stay factored between gas optics and solver (default), 0 = materialised lay/lev_source")
    ap.add_argument("--solver-warps", type=int, default=0, help="solvers (warps) per CTA in the v5 solver kernels (0 = default)")
    ap.add_argument("--lw-solver-gen", type=int, default=0, help="generation of the packed LW solver: 0 default, 5 staged scratch, 6 direct scratch")
    ap.add_argument("--sw-solver-gen", type=int, default=0, help="the same for the SW solver")
    args = ap.parse_args()
    NLAY = args.nlay
    args.steps = max(1, args.steps)
    args.warmup = max(3, args.warmup) if args.impl == "b200" else max(0, args.warmup)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rte_rrtmgp_nn_b200 import api, spectral

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ncol_total = args.columns
    from rte_rrtmgp_nn_b200 import sharding
    c0, c1 = sharding.shard_bounds(ncol_total, rank, world)
    ncol = c1 - c0

    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    ctx.set_flag("fast_math", args.fast_math)
    ctx.set_flag("solver_buffer", args.solver_buffer)
    ctx.set_flag("sw_fast_math", args.sw_fast_math)
    ctx.set_flag("solver_variant", args.solver_variant)
    ctx.set_flag("solver_scratch_mb", args.solver_scratch_mb)
    ctx.set_flag("solver_warps", args.solver_warps)
    ctx.set_flag("lw_solver_gen", args.lw_solver_gen)
    ctx.set_flag("sw_solver_gen", args.sw_solver_gen)
    ctx.set_flag("lw_compact_source", args.lw_compact_source)
    if args.chunk:
        ctx.set_chunk_columns(args.chunk)
    k_lw = api.ty_gas_optics_rrtmgp(ctx); k_lw.load(spectral.synthetic_kdist_lw(NGPT_LW))
    k_sw = api.ty_gas_optics_rrtmgp(ctx); k_sw.load(spectral.synthetic_kdist_sw(NGPT_SW))
    nets_lw = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in LW_FILES]
    nets_sw = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in SW_FILES]

    # ---- inputs: pinned host copies (for e2e) and device-resident copies (for value) ----
    if args.device_inputs:
        args.no_e2e = True
    n_host = min(ncol, NUNIQUE) if args.device_inputs else ncol

    def expand(t):   # device-side tiling to the shard's column count (--device-inputs), identity otherwise
        if t.shape[0] == ncol:
            return t
        reps = -(-ncol // t.shape[0])
        return t.repeat((reps,) + (1,) * (t.dim() - 1))[:ncol].contiguous()

    atm = make_inputs(n_host, NLAY, seed=12345 + rank)
    pin = {}
    for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0"):
        t = torch.empty(atm[k].shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = atm[k]
        pin[k] = t
    gas_pin = {}
    for k, v in atm["gases"].items():
        if np.ndim(v) == 2:
            t = torch.empty(v.shape, dtype=torch.float32, pin_memory=True); t.numpy()[...] = v
            gas_pin[k] = t
    del atm["play"], atm["plev"], atm["tlay"], atm["tlev"]
    with torch.cuda.stream(stream):
        d = {k: expand(v.to(dev, non_blocking=True)) for k, v in pin.items()}
        gas_dev = api.ty_gas_concs()
        gas_host = api.ty_gas_concs()
        for k, v in atm["gases"].items():
            if np.ndim(v) == 2:
                gas_dev.set_vmr(k, expand(gas_pin[k].to(dev, non_blocking=True)))
                gas_host.set_vmr(k, gas_pin[k].numpy())
            else:
                gas_dev.set_vmr(k, float(v)); gas_host.set_vmr(k, float(v))
        nlev = NLAY + 1
        fl = {k: torch.empty((ncol, nlev), dtype=torch.float32, device=dev) for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir")}
        gathered = None
        if world > 1:
            # equal-size slots (the last ranks may hold one column less): gather buffer sized by the largest shard
            nmax = sharding.max_shard(ncol_total, world)
            send = torch.zeros((5, nmax, nlev), dtype=torch.float32, device=dev)
            gathered = torch.empty((world * 5, nmax, nlev), dtype=torch.float32, device=dev)
    stream.synchronize()

    def step_device():
        api.lw_fluxes(k_lw, nets_lw, d["play"], d["plev"], d["tlay"], d["tsfc"], d["sfc_emis"], gas_dev, fl["lw_up"], fl["lw_dn"],
                      tlev=d["tlev"], top_at_1=True, n_gauss_angles=1)
        api.sw_fluxes(k_sw, nets_sw, d["play"], d["plev"], d["tlay"], d["mu0"], d["sfc_alb"], gas_dev, fl["sw_up"], fl["sw_dn"],
                      fl["sw_dir"], top_at_1=True)
        if world > 1:
            with torch.cuda.stream(stream):
                for i, k in enumerate(("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir")):
                    send[i, :ncol].copy_(fl[k])
                dist.all_gather_into_tensor(gathered, send)

    out_host = {} if args.no_e2e else \
        {k: torch.empty((ncol, nlev), dtype=torch.float32, pin_memory=True) for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir")}

    def step_host():
        api.lw_fluxes_host(k_lw, nets_lw, pin["play"].numpy(), pin["plev"].numpy(), pin["tlay"].numpy(), pin["tsfc"].numpy(),
                           pin["sfc_emis"].numpy(), gas_host, tlev=pin["tlev"].numpy(), top_at_1=True, n_gauss_angles=1,
                           flux_up=out_host["lw_up"].numpy(), flux_dn=out_host["lw_dn"].numpy())
        api.sw_fluxes_host(k_sw, nets_sw, pin["play"].numpy(), pin["plev"].numpy(), pin["tlay"].numpy(), pin["mu0"].numpy(),
                           pin["sfc_alb"].numpy(), gas_host, top_at_1=True, flux_up=out_host["sw_up"].numpy(),
                           flux_dn=out_host["sw_dn"].numpy(), flux_dn_dir=out_host["sw_dir"].numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        return ev0.elapsed_time(ev1) / steps, wall / steps * 1e3

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up, then the timed device-resident steps ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count
    ctx.profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(step_device, args.steps)
    if rank == 0:
        sampler.stop_flag = True
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count - launches0
    ms_dev = max_over_ranks(ms_dev)
    value = ncol_total / (ms_dev * 1e-3)

    # ---- end to end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        step_host()  # warm-up (workspace growth)
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        ms_host = (time.perf_counter() - t0) / args.steps * 1e3   # the call is synchronous: wall clock brackets H2D..D2H
        ms_host = max_over_ranks(ms_host)
        lw_in = pin["play"].numel() + pin["plev"].numel() + pin["tlay"].numel() + pin["tlev"].numel() + 2 * ncol
        sw_in = pin["play"].numel() + pin["plev"].numel() + pin["tlay"].numel() + 2 * ncol
        gas2d = sum(t.numel() for t in gas_pin.values())
        h2d = 4 * (lw_in + sw_in + 2 * gas2d)
        d2h = 4 * 5 * ncol * nlev
        e2e = {"value": ncol_total / (ms_host * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_host, "note": "per-rank bytes; wall clock around the synchronous rrnn_{lw,sw}_fluxes_host calls"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peaks()
    abytes = algorithmic_bytes_per_column(NLAY, lw_compact=bool(args.lw_compact_source) and args.solver_variant == 0)
    shares = {k: v[0] for k, v in prof.items()}
    tot = sum(shares.values()) or 1.0
    dom = max(shares, key=shares.get)
    kern = {}
    for k, (ms, n) in prof.items():
        if n:
            gbs = abytes[k] * ncol * args.steps / (ms * 1e-3) / 1e9
            kern[k] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps, "share": ms / tot,
                       "algorithmic_gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak}
    ach = kern[dom]["algorithmic_gb_per_s"]
    nl = kern[dom]["launches_per_step"]
    cols_per_launch = ncol / max(nl, 1.0)
    traffic, traffic_src = ncu_traffic(dom, cols_per_launch)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_column": abytes[dom],
                "algorithmic_bytes_per_launch": abytes[dom] * cols_per_launch, "columns_per_launch": cols_per_launch,
                "ms_per_launch": kern[dom]["ms_per_step"] / max(nl, 1.0),
                "note": "the RTE solvers are bound by instruction issue and latency, not by HBM (ncu, sw_solver: 209 warp instructions per "
                        "64 g-points and layer, issue slots 56 % busy at 12 warps/SM, fp32x2 + MUFU arithmetic); their DRAM traffic exceeds the algorithmic bytes because the reverse-sweep scratch of all resident "
                        "warps is larger than the L2 and partly spills (DESIGN.md section 3)",
                "per_kernel": kern}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, threads, dt = cpu_baseline(args.cpu_columns, NLAY)
        cpu = {"value": v, "unit": "columns/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_columns} columns x {NLAY} layers of the same synthetic workload, LW+SW, {dt:.1f} s; oracle/oracle.c "
                         "(C restatement of the reference kernels, -O3 AVX2 OpenMP over columns) -- the Fortran reference cannot be built here"}

    line = {
        "metric": "LW+SW flux columns/sec (NN gas optics + RTE)", "value": value, "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world, ncol_total),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
        "fast_math": int(args.fast_math), "sw_fast_math": int(args.sw_fast_math), "lw_compact_source": int(bool(args.lw_compact_source) and args.solver_variant == 0), "nn_variant": "tcgen05 (fp16 hi/lo split operands, fp32 accumulation in TMEM)",
        "solver_variant": {0: "v6 TMA-staged packed fp32x2, direct reverse-sweep scratch", 5: "v5 TMA-staged packed fp32x2 (staged scratch)",
                           1: "v3 one g-point per lane"}.get(args.solver_variant, str(args.solver_variant)),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
