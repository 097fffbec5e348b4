#!/usr/bin/env python
"""bench.py -- flux columns/sec of the hot path (NN gas optics + RTE) on N B200s, next to the CPU restatement.

    python bench.py --gpus N --steps K --warmup W                (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --config {gcm,rfmip_lw,rfmip_sw,allsky,sweep} [--models {g256,g128}]

Workloads (BASELINE.json `configs`):
  gcm       configs[3], the headline: synthetic clear-sky LW+SW, 1,000,000 columns x 137 layers, columns sharded contiguously
            over the ranks (strong scaling: the total is fixed)
  rfmip_lw  configs[0]: the 1800 RFMIP columns x 60 layers (real profiles, the drivers' conditioning), LW, 256 g-points
  rfmip_sw  configs[1]: the same columns, SW two-stream + adding, 224 g-points (TSI renormalisation, night columns zeroed)
  allsky    configs[2]: synthetic 100,000 columns x 60 layers, LW+SW with LUT cloud optics, delta-scaling and increment
  sweep     configs[4]: 1e3 ... 1e6 columns x 91 layers LW+SW; one line, `value` = the largest size, `sweep` = every size
One step = one pass of the hot path over all columns of the rank:
    gas_optics(neural_nets=) -> rte_lw      and / or      gas_optics(neural_nets=) -> boundary conditions -> rte_sw
  value : whole-job columns/s with inputs resident in HBM (rrnn_lw_fluxes + rrnn_sw_fluxes on device pointers), timed with
          CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; for N>1 the NCCL
          all-gather of the broadband fluxes is inside the timed step (piecewise, on a side stream under the next piece's kernels).
  e2e   : the same pass through the host-buffer C-ABI calls (rrnn_{lw,sw}_fluxes_host), H2D and D2H copies inside the timed
          region.  `value` is from PAGEABLE caller memory (what a Fortran / C host passes: staged through the library's pinned
          bounce ring), `pinned_value` from page-locked caller memory.
  roofline     : the dominant kernel (largest share of device time), timed live with CUDA events inside the library.
  cpu_baseline : oracle/bench_cpu.c -- compiled C, OpenMP over column blocks, blocked SGEMM over a block's samples, block sizes
                 8 / 36 / 128 / 1800, best of 5, gas-optics / solver split -- on a bounded column sample.
  check        : max |flux - oracle| on 24 sampled columns of the device leg, the e2e leg and (N>1) the gathered buffer.
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
MODELS = {
    "g256": dict(lw=("lw-g256-2018-12-04_absorption_58_58.nc", "lw-g256-2018-12-04_planck_frac_16_16.nc"), ngpt_lw=256,
                 sw=("sw-g224-2018-12-04-absorption_16_16.nc", "sw-g224-2018-12-04-rayleigh_16_16.nc"), ngpt_sw=224,
                 name="g256 (LW 18-58-58-256 + 18-16-16-256) / g224 (SW 7-16-16-224 x2)"),
    "g128": dict(lw=("lw-g128-210809_absorption_BEST.nc", "lw-g128-210809_planck_frac_BEST.nc"), ngpt_lw=128,
                 sw=("sw-g112-210809_absorption_BEST.nc", "sw-g112-210809_rayleigh_BEST.nc"), ngpt_sw=112,
                 name="g128 (LW 18-72-72-128 + 18-24-24-128) / g112 (SW 7-32-32-112 x2), the *_BEST.nc set"),
}
LW_FILES, SW_FILES = MODELS["g256"]["lw"], MODELS["g256"]["sw"]
NN_DIR = os.path.join(ROOT, "data", "nn")
NUNIQUE = 8192  # distinct synthetic columns, tiled to the full size (values do not change the work done)
CONFIGS = {
    "gcm": dict(ncol=NCOL_TOTAL, nlay=137, lw=True, sw=True, clouds=False, data="synthetic",
                what="GCM-scale synthetic clear-sky LW+SW"),
    "rfmip_lw": dict(ncol=1800, nlay=60, lw=True, sw=False, clouds=False, data="rfmip",
                     what="RFMIP clear-sky LW (100 sites x 18 experiments, real profiles)"),
    "rfmip_sw": dict(ncol=1800, nlay=60, lw=False, sw=True, clouds=False, data="rfmip",
                     what="RFMIP clear-sky SW (100 sites x 18 experiments, real profiles)"),
    "allsky": dict(ncol=100_000, nlay=60, lw=True, sw=True, clouds=True, data="synthetic",
                   what="all-sky LW+SW with LUT cloud optics (the example's cloud recipe)"),
    "sweep": dict(ncol=1_000_000, nlay=91, lw=True, sw=True, clouds=False, data="synthetic",
                  what="column-count sweep, clear-sky LW+SW", sizes=(1_000, 10_000, 100_000, 1_000_000)),
}
# the reference's own published timings (ifort + MKL, ONE core, 1800 columns x 60 layers, NN gas optics):
# LW 84.4 + ~3 + 99.0 ms, SW 42.0 + ~24 + 229.0 ms
PUBLISHED_MS_1800x60 = {"lw": 186.0, "sw": 295.0}


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


def rfmip_inputs():
    """The 1800 RFMIP columns with the drivers' conditioning (rte_rrtmgp_nn_b200.rfmip); night columns carry mu0 <= 0."""
    from rte_rrtmgp_nn_b200 import rfmip
    atm = rfmip.load()
    atm["mu0_driver"] = np.where(atm["usecol"], atm["mu0"], -1.0).astype(np.float32)
    return atm


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  ONE long-lived `nvidia-smi -lms` process is started
    when the sampler is constructed (before the warm-up): forking a helper from a process that holds gigabytes of pinned memory
    inside the timed region stalls the launching thread for milliseconds (measured: 5 - 12 ms lost per step at 100 000 columns
    whenever the fork fell into the region).  start() / stop_flag only mark which of the samples count."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_ms=100):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.all = []          # (time, fields) of every line the helper printed
        self.stop_flag = False
        self.t0 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index),
                                          "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.reader = threading.Thread(target=self._read, daemon=True)
            self.reader.start()
            t_end = time.perf_counter() + 5.0      # the helper's NVML start-up (~0.5 s) contends for driver locks: wait it out here
            while not self.all and time.perf_counter() < t_end and self.proc.poll() is None:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        try:
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 7:
                    self.all.append((time.perf_counter(), f))
        except Exception:
            pass

    def run(self):
        self.t0 = time.perf_counter()
        while not self.stop_flag:
            time.sleep(0.005)
        t1 = time.perf_counter()
        inside = [f for t, f in list(self.all) if self.t0 <= t <= t1 + 0.05]
        if not inside and self.all:    # a timed region shorter than the sampling period: the sample nearest to it
            inside = [min(list(self.all), key=lambda tf: abs(tf[0] - 0.5 * (self.t0 + t1)))[1]]
        self.samples = inside
        self.close()

    def close(self):
        if self.proc is not None:
            try:
                self.proc.kill()
            except Exception:
                pass
            self.proc = None

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


def algorithmic_bytes_per_column(nlay, G=NGPT_LW, H=NGPT_SW, nx_lw=18, nx_sw=7, lw_compact=True):
    """bytes each kernel must move per column (fp32), every array touched once.
    lw_compact: the LW sources cross HBM factored -- tau, pfrac (G,L) and two band tables (16,L), (16,L+1) instead of
    tau, lay_source (G,L) and lev_source (G,L+1); the denominators shrink with the traffic (8 instead of 12 B per g-point
    and layer), so `frac` stays a statement about the bytes that really have to move."""
    L = nlay
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


# ------------------------------------------------------------------------------------------------ CPU side (oracle; test infrastructure)
def oracle_nets(models):
    import nc4min
    import oracle as O
    m = MODELS[models]
    return ([O.Net(nc4min.load_nn_model(os.path.join(NN_DIR, f))) for f in m["lw"]],
            [O.Net(nc4min.load_nn_model(os.path.join(NN_DIR, f))) for f in m["sw"]])


def cpu_problem(ncol_sample, nlay, models="g256", data="synthetic"):
    """The compiled CPU driver's copy of the workload (oracle/bench_cpu.py)."""
    import bench_cpu
    from rte_rrtmgp_nn_b200 import spectral
    m = MODELS[models]
    atm = rfmip_inputs() if data == "rfmip" else make_inputs(ncol_sample, nlay, seed=777)  # (night columns: mu0 = 1, as the reference driver)
    lw, sw = oracle_nets(models)
    return bench_cpu.Problem(spectral.synthetic_kdist_lw(m["ngpt_lw"]), spectral.synthetic_kdist_sw(m["ngpt_sw"]), lw, sw, atm)


def cpu_baseline(ncol_sample, nlay, models="g256", lw=True, sw=True, repeats=5, data="synthetic", blocks=(8, 36, 128, 1800)):
    """block sizes 8 / 36 / 128 / 1800, best of `repeats`, OpenMP over blocks, gas-optics / solver split."""
    P = cpu_problem(ncol_sample, nlay, models, data)
    t0 = time.perf_counter()
    sweep = [P.run(b, repeats=repeats, lw=lw, sw=sw) for b in blocks if b <= max(P.ncol, 8)]
    best = max(sweep, key=lambda r: r["columns_per_s"])
    threads = best["threads"]
    pub = sum(PUBLISHED_MS_1800x60[k] for k, on in (("lw", lw), ("sw", sw)) if on)
    pub_per_core = 1800.0 / (pub * 1e-3) * 60.0 / nlay
    band = ("LW" if lw else "") + ("+" if lw and sw else "") + ("SW" if sw else "")
    return {
        "value": best["columns_per_s"], "unit": "columns/s", "cores": threads, "kind": "port",
        "sample": f"{P.ncol} columns x {nlay} layers of the same workload ({band}), best of {repeats} passes per block size, "
                  f"{time.perf_counter() - t0:.1f} s in all; oracle/bench_cpu.c: compiled C (-Ofast, AVX2/FMA), OpenMP over column blocks, 6x16 "
                  "register-blocked SGEMM over a block's nlay*block samples, everything else oracle.c's restatement of the reference kernels "
                  "-- the Fortran reference cannot be built here (no Fortran compiler in the image)",
        "best_block_size": best["block"],
        "block_size_sweep": [{"block": r["block"], "columns_per_s": r["columns_per_s"], "gas_optics_s_per_thread": r["gas_optics_s_per_thread"],
                              "solver_s_per_thread": r["solver_s_per_thread"]} for r in sweep],
        "gas_optics_share": best["gas_optics_s_per_thread"] / max(best["gas_optics_s_per_thread"] + best["solver_s_per_thread"], 1e-30),
        "columns_per_s_per_core": best["columns_per_s"] / threads,
        "published_reference_columns_per_s_per_core": pub_per_core,
        "published_note": f"BASELINE.md section 1: the reference's own figure, ifort + MKL on ONE core, {pub:.0f} ms per 1800 columns x 60 layers "
                          f"(NN gas optics + solver), scaled linearly to {nlay} layers; this C port reaches "
                          f"{best['columns_per_s'] / threads / pub_per_core:.2f} of it per core",
    }, P


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference is Fortran and cannot be
    built in this image (no Fortran compiler), so this is the compiled C port (oracle/bench_cpu.c) with all
    host threads: each step is one pass over a bounded column SAMPLE of the named workload (throughput is linear in columns)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is entitled to every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cfg = CONFIGS[args.config]
    nlay = args.nlay or cfg["nlay"]
    ncol_sample = 1800 if cfg["data"] == "rfmip" else min(args.cpu_columns, cfg["ncol"])
    base, P = cpu_baseline(ncol_sample, nlay, args.models, cfg["lw"], cfg["sw"], repeats=1, data=cfg["data"])
    vals = []
    for i in range(args.warmup + args.steps):
        r = P.run(base["best_block_size"], repeats=1, lw=cfg["lw"], sw=cfg["sw"])
        if i >= args.warmup:
            vals.append(r)
    value = float(np.mean([r["columns_per_s"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
    base["value"] = value
    base["sample"] = (f"each step = ONE pass over a {P.ncol}-column SAMPLE of the workload (not all {cfg['ncol']} columns: throughput is "
                      f"linear in columns) at the best block size ({base['best_block_size']}); " + base["sample"])
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": cfg["data"],
        "config": workload_config(args.config, args.models, args.gpus, cfg["ncol"], nlay),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def metric_name(cfg):
    band = "LW+SW" if cfg["lw"] and cfg["sw"] else ("LW" if cfg["lw"] else "SW")
    return f"{band} flux columns/sec (NN gas optics + RTE)"


def workload_config(config, models, ngpus, ncol_total, nlay):
    cfg, m = CONFIGS[config], MODELS[models]
    return {"workload": f"{cfg['what']}: {ncol_total} columns x {nlay} layers, NN gas optics {m['name']}"
                        + (" + rte_lw (1 angle)" if cfg["lw"] else "") + (" + rte_sw (two-stream)" if cfg["sw"] else "")
                        + (" + cloud optics (LUT), delta-scaling, increment" if cfg["clouds"] else ""),
            "name": config, "models": models, "ncol_total": ncol_total, "nlay": nlay,
            "ngpt_lw": m["ngpt_lw"] if cfg["lw"] else None, "ngpt_sw": m["ngpt_sw"] if cfg["sw"] else None,
            "sharding": f"columns split contiguously over {ngpus} rank(s)",
            "l2": "per-step working set (optical properties, 0.4 - 1.6 MB per column) is far larger than the 126 MB L2 at every size but the "
                  "1000-column point of the sweep; no explicit flush"}


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


# ------------------------------------------------------------------------------------------------ the check against the oracle
def oracle_fluxes(cfg, models, atm, idx, clouds=None, cloud_tables=None):
    """Strict fp32 oracle on the columns `idx` of the workload -> dict of (len(idx), nlev) arrays."""
    import oracle as O
    from rte_rrtmgp_nn_b200 import spectral
    m = MODELS[models]
    lw_n, sw_n = oracle_nets(models)
    n0 = atm["play"].shape[0]
    sub = {k: np.ascontiguousarray(v[idx]) for k, v in atm.items() if isinstance(v, np.ndarray) and v.shape[:1] == (n0,)}
    g = {k: (np.ascontiguousarray(v[idx]) if np.ndim(v) == 2 else v) for k, v in atm["gases"].items()}
    top = bool(atm.get("top_at_1", True))
    out = {}
    if cfg["lw"]:
        kd = spectral.synthetic_kdist_lw(m["ngpt_lw"])
        r = O.gas_optics_lw(kd, lw_n, sub["play"], sub["plev"], sub["tlay"], sub["tsfc"], g, tlev=sub["tlev"])
        tau = r["tau"]
        if clouds is not None:
            ctau = O.cloud_optics_lut(cloud_tables["lw"], *[clouds[k][idx] for k in ("lwp", "iwp", "rel", "rei")], False)
            tau = O.inc_1scalar_by_1scalar_bybnd(tau, ctau, kd["band_lims_gpt"])
        emis = np.repeat(sub["sfc_emis"][:, None], kd["nbnd"], 1)
        out["lw_up"], out["lw_dn"] = O.rte_lw(kd, top, tau, r["lay_source"], r["lev_source"], r["sfc_source"], emis)
    if cfg["sw"]:
        ks = spectral.synthetic_kdist_sw(m["ngpt_sw"])
        mu0_in = sub.get("mu0_driver", sub["mu0"])
        night = ~(mu0_in > 0)
        mu0 = np.where(night, np.float32(1.0), mu0_in).astype(np.float32)
        r = O.gas_optics_sw(ks, sw_n, sub["play"], sub["plev"], sub["tlay"], g)
        tau, ssa, gg = r["tau"], r["ssa"], r["g"]
        if clouds is not None:
            c = O.delta_scale_2str(*O.cloud_optics_lut(cloud_tables["sw"], *[clouds[k][idx] for k in ("lwp", "iwp", "rel", "rei")], True))
            tau, ssa, gg = O.inc_2stream_by_2stream_bybnd(tau, ssa, gg, *c, ks["band_lims_gpt"])
        toa = r["toa_src"]
        if "tsi" in sub:   # rrtmgp_rfmip_sw.F90:409-416
            def_tsi = np.float32(0.0)
            for v in np.asarray(ks["solar_source"], np.float32):
                def_tsi = np.float32(def_tsi + v)
            toa = (toa * sub["tsi"][:, None] / def_tsi).astype(np.float32)
        alb = np.repeat(sub["sfc_alb"][:, None], m["ngpt_sw"], 1)
        up, dn, dr = O.rte_sw(top, mu0, toa, alb, alb, tau, ssa, gg)
        up[night] = 0.0; dn[night] = 0.0                      # :458-463
        out["sw_up"], out["sw_dn"], out["sw_dir"] = up, dn, dr
    return out


def max_diff(got, want):
    return {k: float(np.abs(np.asarray(got[k], np.float64) - want[k]).max()) for k in want}


# ------------------------------------------------------------------------------------------------ the GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from rte_rrtmgp_nn_b200 import api, spectral, sharding, synth

    cfg = CONFIGS[args.config]
    m = MODELS[args.models]
    nlay = args.nlay or cfg["nlay"]
    do_lw, do_sw, cloudy = cfg["lw"], cfg["sw"], cfg["clouds"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The flux gather overlaps the next piece's kernels (below).  Optional cap on NCCL's CTAs (an experiment that lost; the
        # solvers instead tolerate a co-running kernel by taking their columns dynamically, rte_solvers_tma.cu: NextColumns).
        opts = None
        if args.nccl_max_ctas > 0:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(args.nccl_max_ctas)
                opts.config.min_ctas = 1
            except Exception:
                opts = None
        if opts is not None:
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=dev)

    stream = torch.cuda.Stream(device=dev)
    side = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    for k in ("fast_math", "solver_buffer", "sw_fast_math", "solver_variant", "solver_scratch_mb", "solver_warps", "lw_compact_source"):
        ctx.set_flag(k, getattr(args, k))
    if world > 1:   # the ranks of one node share its host cores: the threads that stage pageable caller memory are divided among them
        ctx.set_flag("host_copy_threads", max(2, min(8, (os.cpu_count() or 16) // world)))
    if args.chunk:
        ctx.set_chunk_columns(args.chunk)
    k_lw = api.ty_gas_optics_rrtmgp(ctx); k_lw.load(spectral.synthetic_kdist_lw(m["ngpt_lw"]))
    k_sw = api.ty_gas_optics_rrtmgp(ctx); k_sw.load(spectral.synthetic_kdist_sw(m["ngpt_sw"]))
    nets_lw = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in m["lw"]]
    nets_sw = [api.rrtmgp_network_type(ctx).load_netcdf(os.path.join(NN_DIR, f)) for f in m["sw"]]
    nlev = nlay + 1
    names = (["lw_up", "lw_dn"] if do_lw else []) + (["sw_up", "sw_dn", "sw_dir"] if do_sw else [])
    in_keys = ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis", "sfc_alb", "mu0") + (("tsi",) if cfg["data"] == "rfmip" else ())

    def one_size(ncol_total, with_e2e, with_check):
        """Everything for one column count: inputs, the timed device leg, the e2e legs, the check.  Returns a dict."""
        c0, c1 = sharding.shard_bounds(ncol_total, rank, world)
        ncol = c1 - c0
        device_inputs = args.device_inputs
        n_host = min(ncol, NUNIQUE) if device_inputs else ncol
        if cfg["data"] == "rfmip":
            full = rfmip_inputs()
            atm = {k: (v[c0:c1] if isinstance(v, np.ndarray) else v) for k, v in full.items()}
            atm["gases"] = {k: v[c0:c1] for k, v in full["gases"].items()}
            atm["mu0"] = atm["mu0_driver"]
        else:
            atm = make_inputs(n_host, nlay, seed=12345 + rank)
        top = bool(atm.get("top_at_1", True))
        clouds = synth.make_clouds(atm) if cloudy else None

        def expand(t):   # device-side tiling to the shard's column count (--device-inputs), identity otherwise
            if t.shape[0] == ncol:
                return t
            reps = -(-ncol // t.shape[0])
            return t.repeat((reps,) + (1,) * (t.dim() - 1))[:ncol].contiguous()

        # ---- inputs: pinned host copies, plain (pageable) numpy copies (both for e2e) and device-resident copies (for value)
        pin, gas_pin = {}, {}
        for k in in_keys:
            t = torch.empty(atm[k].shape, dtype=torch.float32, pin_memory=True)
            t.numpy()[...] = atm[k]
            pin[k] = t
        for k, v in atm["gases"].items():
            if np.ndim(v) == 2:
                t = torch.empty(v.shape, dtype=torch.float32, pin_memory=True); t.numpy()[...] = v
                gas_pin[k] = t
        with torch.cuda.stream(stream):
            d = {k: expand(v.to(dev, non_blocking=True)) for k, v in pin.items()}
            gas_dev, gas_host, gas_page = api.ty_gas_concs(), api.ty_gas_concs(), api.ty_gas_concs()
            for k, v in atm["gases"].items():
                if np.ndim(v) == 2:
                    gas_dev.set_vmr(k, expand(gas_pin[k].to(dev, non_blocking=True)))
                    gas_host.set_vmr(k, gas_pin[k].numpy())
                    gas_page.set_vmr(k, np.array(v, np.float32, copy=True))
                else:
                    for gc in (gas_dev, gas_host, gas_page):
                        gc.set_vmr(k, float(v))
            cl_dev = {k: expand(torch.from_numpy(v).to(dev)) for k, v in clouds.items()} if cloudy else None
        stream.synchronize()
        tsi_d = d.get("tsi")

        # ---- output buffers.  N > 1: the rank's fluxes are written straight into its slot of the gather buffers, one buffer
        # per piece of the shard, each all-gathered in place on a side stream while the next piece is being computed.
        npiece = 1 if world == 1 else max(1, min(args.gather_pieces, ncol // 4096))
        bounds = [(ncol * i // npiece, ncol * (i + 1) // npiece) for i in range(npiece)]

        def piece_len(r, i):
            n = sharding.shard_bounds(ncol_total, r, world)[1] - sharding.shard_bounds(ncol_total, r, world)[0]
            return n * (i + 1) // npiece - n * i // npiece
        nmaxp = [max(piece_len(r, i) for r in range(world)) for i in range(npiece)]
        nf = len(names)
        if world > 1:
            gbuf = [torch.zeros((world, nf, nmaxp[i], nlev), dtype=torch.float32, device=dev) for i in range(npiece)]
            outs = [{nm: gbuf[i][rank, j, :b - a] for j, nm in enumerate(names)} for i, (a, b) in enumerate(bounds)]
        else:
            gbuf = None
            outs = [{nm: torch.empty((ncol, nlev), dtype=torch.float32, device=dev) for nm in names}]
        gas_piece = []
        for (a, b) in bounds:
            gc = api.ty_gas_concs()
            for k, v in atm["gases"].items():
                gc.set_vmr(k, gas_dev.get_vmr(k)[a:b] if np.ndim(v) == 2 else float(v))
            gas_piece.append(gc)
        # The gather itself.  "nccl" (default): in-place all_gather_into_tensor per piece on a side stream, NCCL held to
        # --nccl-max-ctas CTAs so that its kernel leaves the SMs to the persistent solver clusters of the next piece.
        # "p2p" (experiment, kept): every rank pushes its slot of a piece into every peer's gather buffer with device-to-device copies
        # through CUDA IPC peer mappings (copy engines, no SM at all; one 4-byte all_reduce after the last piece as the completion
        # signal).  Correct, and the solvers run undisturbed -- but torch's cross-process copy reached only 26 GB/s on the 2-GPU box
        # (345 MB per piece in 13 ms; profiles/r2o_n2_p2p.json), so the last piece's copies are exposed: 217 against 208 ms per pass.
        peer = None
        gather_mode = args.gather if world > 1 else "none"
        if gather_mode == "p2p":
            try:
                from torch.multiprocessing.reductions import reduce_tensor
                mine = [reduce_tensor(g) for g in gbuf]
                everyone = [None] * world
                dist.all_gather_object(everyone, mine)
                peer = [[g if r == rank else fn(*a) for g, (fn, a) in zip(gbuf, everyone[r])] for r in range(world)]
                for r in range(world):     # first touch of every peer mapping outside the timed region
                    if r != rank:
                        peer[r][0][rank, 0, :1].copy_(gbuf[0][rank, 0, :1])
                torch.cuda.synchronize()
                dist.barrier()
            except Exception as e:   # no peer access / IPC on this box: NCCL does it
                print(f"[bench] rank {rank}: p2p gather unavailable ({e}); using NCCL all_gather", file=sys.stderr)
                peer, gather_mode = None, "nccl"
            flag = torch.ones(1, dtype=torch.float32, device=dev)
            modes = [None] * world
            dist.all_gather_object(modes, gather_mode)
            if any(mm != "p2p" for mm in modes):
                peer, gather_mode = None, "nccl"
        ev_piece = [torch.cuda.Event() for _ in bounds]
        ev_g0 = [torch.cuda.Event(enable_timing=True) for _ in bounds]
        ev_g1 = [torch.cuda.Event(enable_timing=True) for _ in bounds]
        ev_cmp_end = torch.cuda.Event(enable_timing=True)

        # ---- all-sky (configs[2]): examples/all-sky/rrtmgp_allsky.F90:366-446 for all columns -- cloud optics (LUT) by band,
        # gas optics, delta-scaling, increment, rte -- through the fused drivers (rrnn_{lw,sw}_fluxes_allsky[_host]): the cloud
        # increment happens inside the solvers
        allsky = None
        if cloudy:
            def lut(band):
                return api.load_cloud_lut_file(os.path.join(ROOT, "data", "cloud_optics", f"rrtmgp-cloud-optics-coeffs-{band}.nc"))
            co_lw = api.ty_cloud_optics(ctx); assert co_lw.load(**lut("lw")) == ""
            co_sw = api.ty_cloud_optics(ctx); assert co_sw.load(**lut("sw")) == ""
            allsky = dict(co_lw=co_lw, co_sw=co_sw)

        def ok(msg):
            if msg != "":
                raise RuntimeError(msg)

        def allsky_pass(src_in, gases, cl, out):
            """One pass over the shard; src_in / cl: device tensors of the whole shard."""
            A = allsky
            with torch.cuda.stream(stream):
                api.lw_fluxes_allsky(k_lw, nets_lw, A["co_lw"], src_in["play"], src_in["plev"], src_in["tlay"], src_in["tsfc"], src_in["sfc_emis"],
                                     gases, cl, out["lw_up"], out["lw_dn"], tlev=src_in["tlev"], top_at_1=top)
                api.sw_fluxes_allsky(k_sw, nets_sw, A["co_sw"], src_in["play"], src_in["plev"], src_in["tlay"], src_in["mu0"], src_in["sfc_alb"],
                                     gases, cl, out["sw_up"], out["sw_dn"], out["sw_dir"], top_at_1=top)

        def step_device():
            if cloudy:
                allsky_pass(d, gas_dev, cl_dev, outs[0])
                return
            for i, (a, b) in enumerate(bounds):
                o = outs[i]
                if do_lw:
                    api.lw_fluxes(k_lw, nets_lw, d["play"][a:b], d["plev"][a:b], d["tlay"][a:b], d["tsfc"][a:b], d["sfc_emis"][a:b], gas_piece[i],
                                  o["lw_up"], o["lw_dn"], tlev=d["tlev"][a:b], top_at_1=top, n_gauss_angles=1)
                if do_sw:
                    api.sw_fluxes(k_sw, nets_sw, d["play"][a:b], d["plev"][a:b], d["tlay"][a:b], d["mu0"][a:b], d["sfc_alb"][a:b], gas_piece[i],
                                  o["sw_up"], o["sw_dn"], o["sw_dir"], tsi=None if tsi_d is None else tsi_d[a:b], top_at_1=top)
                if world > 1:
                    ev_piece[i].record(stream)
                    with torch.cuda.stream(side):
                        side.wait_event(ev_piece[i])
                        ev_g0[i].record(side)
                        if peer is not None:
                            for k in range(1, world):          # staggered targets: no two ranks push to the same peer at once
                                r = (rank + k) % world
                                peer[r][i][rank].copy_(gbuf[i][rank], non_blocking=True)
                            if i == npiece - 1:
                                dist.all_reduce(flag)           # ordered after this rank's pushes: completion = everybody's have landed
                        else:
                            dist.all_gather_into_tensor(gbuf[i].view(world * nf, nmaxp[i], nlev), gbuf[i][rank])   # in place
                        ev_g1[i].record(side)
            if world > 1:
                ev_cmp_end.record(stream)
                stream.wait_stream(side)

        # ---- e2e: host buffers in, host buffers out, through the C ABI (or, for all-sky, the same API with host tensors)
        page = {k: np.array(atm[k], np.float32, copy=True) for k in in_keys}   # plain numpy: pageable memory
        out_pin = {nm: torch.empty((ncol, nlev), dtype=torch.float32, pin_memory=True) for nm in names} if with_e2e else {}
        out_page = {nm: np.empty((ncol, nlev), np.float32) for nm in names} if with_e2e else {}

        def as_np(x):
            return x if isinstance(x, np.ndarray) else x.numpy()

        def step_host(src, gases, out):
            if cloudy:
                A = allsky
                api.lw_fluxes_allsky_host(k_lw, nets_lw, A["co_lw"], as_np(src["play"]), as_np(src["plev"]), as_np(src["tlay"]), as_np(src["tsfc"]),
                                          as_np(src["sfc_emis"]), gases, clouds, tlev=as_np(src["tlev"]), top_at_1=top,
                                          flux_up=as_np(out["lw_up"]), flux_dn=as_np(out["lw_dn"]))
                api.sw_fluxes_allsky_host(k_sw, nets_sw, A["co_sw"], as_np(src["play"]), as_np(src["plev"]), as_np(src["tlay"]), as_np(src["mu0"]),
                                          as_np(src["sfc_alb"]), gases, clouds, top_at_1=top, flux_up=as_np(out["sw_up"]),
                                          flux_dn=as_np(out["sw_dn"]), flux_dn_dir=as_np(out["sw_dir"]))
                return
            if do_lw:
                api.lw_fluxes_host(k_lw, nets_lw, as_np(src["play"]), as_np(src["plev"]), as_np(src["tlay"]), as_np(src["tsfc"]),
                                   as_np(src["sfc_emis"]), gases, tlev=as_np(src["tlev"]), top_at_1=top, n_gauss_angles=1,
                                   flux_up=as_np(out["lw_up"]), flux_dn=as_np(out["lw_dn"]))
            if do_sw:
                api.sw_fluxes_host(k_sw, nets_sw, as_np(src["play"]), as_np(src["plev"]), as_np(src["tlay"]), as_np(src["mu0"]),
                                   as_np(src["sfc_alb"]), gases, tsi=as_np(src["tsi"]) if "tsi" in src else None, top_at_1=top,
                                   flux_up=as_np(out["sw_up"]), flux_dn=as_np(out["sw_dn"]), flux_dn_dir=as_np(out["sw_dir"]))

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps):
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record(stream)
            for _ in range(steps):
                fn()
            ev1.record(stream)
            barrier()
            return ev0.elapsed_time(ev1) / steps

        def max_over_ranks(ms):
            if world == 1:
                return ms
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # ---- warm-up, then the timed device-resident steps ----
        sampler = ClockSampler(local_rank) if rank == 0 else None      # (its helper process is forked here, outside the timed region)
        for _ in range(args.warmup):
            step_device()
        barrier()
        launches0 = ctx.launch_count
        tc0, ff0 = ctx.nn_kernel_counts
        ctx.profile(True)
        if rank == 0:
            sampler.start()
        ms_dev = timed(step_device, args.steps)
        if rank == 0:
            sampler.stop_flag = True
            sampler.join(timeout=2.0)
        prof = ctx.profile_read()
        ctx.profile(False)
        launches = ctx.launch_count - launches0
        tc1, ff1 = ctx.nn_kernel_counts
        ms_dev = max_over_ranks(ms_dev)
        res = dict(ncol=ncol, ncol_total=ncol_total, ms_dev=ms_dev, value=ncol_total / (ms_dev * 1e-3), prof=prof, launches=launches,
                   clocks=sampler.summary() if sampler is not None else None, nn_kernels={"tcgen05": tc1 - tc0, "fp32_ffma": ff1 - ff0})
        if world > 1:
            res["nccl"] = {"collectives_per_step": npiece if peer is None else 1, "gather": gather_mode,
                           "op": ("all_gather_into_tensor (in place) of the rank's flux slots, per piece of the shard, on a side stream" if peer is None else
                                  "per piece of the shard: device-to-device copies of the rank's flux slots into every peer's gather buffer (CUDA IPC "
                                  "peer memory over NVLink, copy engines, side stream); one 4-byte NCCL all_reduce after the last piece as the completion signal"),
                           "bytes_received_per_rank_per_step": int(sum(4 * (world - 1) * nf * n * nlev for n in nmaxp)),
                           "ms_sum_of_collectives": float(sum(a.elapsed_time(b) for a, b in zip(ev_g0, ev_g1))),
                           "ms_exposed_after_last_kernel": float(max(0.0, ev_cmp_end.elapsed_time(ev_g1[-1]))),
                           "note": "last step of the timed region, rank 0; a collective's time on the side stream includes waiting for the slowest rank"}

        # ---- end to end through the host-buffer C ABI: page-locked and pageable caller memory ----
        if with_e2e:
            e2e = {}
            for label, src, gases, out in (("pinned", pin, gas_host, out_pin), ("pageable", page, gas_page, out_page)):
                step_host(src, gases, out)  # warm-up (workspace / bounce-ring growth)
                step_host(src, gases, out)
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step_host(src, gases, out)
                barrier()
                e2e[label] = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)   # synchronous calls: wall clock brackets H2D..D2H
            per = lambda k: int(np.prod(atm[k].shape[1:]))
            gas2d_n = sum(int(np.prod(v.shape[1:])) for v in atm["gases"].values() if np.ndim(v) == 2)
            n_in = 0
            if True:
                if do_lw:
                    n_in += sum(per(k) for k in ("play", "plev", "tlay", "tlev", "tsfc", "sfc_emis")) + gas2d_n + (4 * nlay if cloudy else 0)
                if do_sw:
                    n_in += sum(per(k) for k in ("play", "plev", "tlay", "mu0", "sfc_alb")) + (1 if "tsi" in in_keys else 0) + gas2d_n + (4 * nlay if cloudy else 0)
            res["e2e"] = {"value": ncol_total / (e2e["pinned"] * 1e-3), "unit": "columns/s",
                          "h2d_bytes_per_step": int(4 * n_in * ncol), "d2h_bytes_per_step": int(4 * nf * ncol * nlev),
                          "ms_per_step": e2e["pinned"],
                          "host_memory": "pinned (page-locked caller buffers, the base contract's definition); pageable_value: the same calls on "
                                         "plain numpy arrays -- what a Fortran host's allocate()d arrays are -- which the library stages through "
                                         "its pinned bounce ring with a few host threads (bounded by host memory bandwidth when 8 ranks share a node)",
                          "pageable_value": ncol_total / (e2e["pageable"] * 1e-3), "pageable_ms_per_step": e2e["pageable"],
                          "pinned_value": ncol_total / (e2e["pinned"] * 1e-3), "pinned_ms_per_step": e2e["pinned"],
                          "note": "per-rank bytes; wall clock around the synchronous host-buffer calls"}

        # ---- check: sampled columns against the strict fp32 oracle (test infrastructure used as the checker only) ----
        if with_check and rank == 0:
            idx = np.unique(np.linspace(0, ncol - 1, min(24, ncol)).astype(np.int64))
            ridx = idx % n_host if (device_inputs and ncol > n_host) else idx
            tabs = {"lw": allsky["co_lw"].tables, "sw": allsky["co_sw"].tables} if cloudy else None
            want = oracle_fluxes(cfg, args.models, atm, ridx, clouds, tabs)
            chk = {"columns_checked": int(len(idx)), "reference": "oracle/oracle.c, strict fp32 build", "unit": "W m-2"}
            torch.cuda.synchronize()
            tidx = torch.from_numpy(idx).to(dev)
            full = {nm: torch.cat([o[nm] for o in outs], 0) for nm in names}   # this rank's own columns (N > 1: its slots of the gather buffers)
            chk["device_max_abs_diff"] = max_diff({nm: full[nm][tidx].cpu().numpy() for nm in names}, want)
            if with_e2e:
                chk["e2e_pageable_max_abs_diff"] = max_diff({nm: out_page[nm][idx] for nm in names}, want)
                chk["e2e_pinned_max_abs_diff"] = max_diff({nm: out_pin[nm].numpy()[idx] for nm in names}, want)
            if world > 1:
                # the gathered buffer: the LAST rank's shard as rank 0 received it, against the oracle on that rank's inputs
                r = world - 1
                rc0, rc1 = sharding.shard_bounds(ncol_total, r, world)
                nr = rc1 - rc0
                atm_r = make_inputs(nr, nlay, seed=12345 + r)
                idr = np.unique(np.linspace(0, nr - 1, 12).astype(np.int64))
                want_r = oracle_fluxes(cfg, args.models, atm_r, idr)
                rb = [(nr * i // npiece, nr * (i + 1) // npiece) for i in range(npiece)]
                tr = torch.from_numpy(idr).to(dev)
                got_r = {nm: torch.cat([gbuf[i][r, j, :b - a] for i, (a, b) in enumerate(rb)], 0)[tr].cpu().numpy() for j, nm in enumerate(names)}
                chk["gathered_last_rank_max_abs_diff"] = max_diff(got_r, want_r)
            chk["checksum_sum_of_fluxes"] = {nm: float(full[nm].double().sum().item()) for nm in names}
            res["check"] = chk
        return res

    # ---------------------------------------------------------------------------------------------- run
    if args.config == "sweep":
        rows = []
        for n in cfg["sizes"][:-1]:
            r = one_size(int(n), with_e2e=not args.no_e2e, with_check=False)
            rows.append({"ncol_total": int(n), "columns_per_s": r["value"], "ms_per_step": r["ms_dev"],
                         "e2e_columns_per_s": r.get("e2e", {}).get("value"), "e2e_pinned_columns_per_s": r.get("e2e", {}).get("pinned_value")})
        ncol_total = int(cfg["sizes"][-1])
        res = one_size(ncol_total, with_e2e=not args.no_e2e, with_check=not args.no_check)
        rows.append({"ncol_total": ncol_total, "columns_per_s": res["value"], "ms_per_step": res["ms_dev"],
                     "e2e_columns_per_s": res.get("e2e", {}).get("value"), "e2e_pinned_columns_per_s": res.get("e2e", {}).get("pinned_value")})
        res["sweep"] = rows
    else:
        ncol_total = args.columns or cfg["ncol"]
        res = one_size(ncol_total, with_e2e=not (args.no_e2e or args.device_inputs), with_check=not args.no_check)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peaks()
    compact = bool(args.lw_compact_source) and args.solver_variant == 0
    abytes = algorithmic_bytes_per_column(nlay, m["ngpt_lw"], m["ngpt_sw"], lw_compact=compact)
    if cloudy:   # clouds folded into the solvers: the solvers read 64 (LW) / 192 (SW) more bytes per layer and column
        abytes["lw_solver"] += 64 * nlay
        abytes["sw_solver"] += 192 * nlay
    prof, ncol = res["prof"], res["ncol"]
    tot = sum(v[0] for v in prof.values()) or 1.0
    kern = {}
    for k, (ms, n) in prof.items():
        if n:
            gbs = abytes[k] * ncol * args.steps / (ms * 1e-3) / 1e9
            kern[k] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps, "share": ms / tot,
                       "algorithmic_gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak}
    dom = max(kern, key=lambda k: kern[k]["share"])
    ach = kern[dom]["algorithmic_gb_per_s"]
    nl = kern[dom]["launches_per_step"]
    cols_per_launch = ncol / max(nl, 1.0)
    traffic, traffic_src = ncu_traffic(dom, cols_per_launch) if (args.models == "g256" and nlay == 137) else (None, None)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_column": abytes[dom],
                "algorithmic_bytes_per_launch": abytes[dom] * cols_per_launch, "columns_per_launch": cols_per_launch,
                "ms_per_launch": kern[dom]["ms_per_step"] / max(nl, 1.0),
                "note": "the RTE solvers are bound by instruction issue and latency, not by HBM (profiles/r2_sw_solver_ablation.md: ~145 warp "
                        "instructions per 64 g-points and layer at 0.53 issued per cycle and scheduler; time follows the instruction count, not "
                        "the bytes); the roofline line is reported against HBM as the contract asks",
                "per_kernel": kern}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ncs = 1800 if cfg["data"] == "rfmip" else min(args.cpu_columns, cfg["ncol"])
        cpu, _ = cpu_baseline(ncs, nlay, args.models, do_lw, do_sw, repeats=5, data=cfg["data"])

    line = {
        "metric": metric_name(cfg), "value": res["value"], "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_dev"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": cfg["data"], "config": workload_config(args.config, args.models, world, ncol_total, nlay),
        "e2e": res.get("e2e"), "gpu_launches": int(res["launches"]), "roofline": roofline, "cpu_baseline": cpu, "clocks": res["clocks"],
        "check": res.get("check"), "nn_kernel_launches": res["nn_kernels"],
        "fast_math": int(args.fast_math), "sw_fast_math": int(args.sw_fast_math), "lw_compact_source": int(compact),
        "nn_variant": "tcgen05 (fp16 hi/lo split operands, fp32 accumulation in TMEM)" if res["nn_kernels"]["fp32_ffma"] == 0 else "fp32 FFMA (fallback)",
        "solver_variant": {0: "v6 TMA-staged packed fp32x2", 1: "v3 one g-point per lane"}.get(args.solver_variant, str(args.solver_variant)),
    }
    if "nccl" in res:
        line["nccl"] = res["nccl"]
    if "sweep" in res:
        line["sweep"] = res["sweep"]
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="gcm", choices=sorted(CONFIGS), help="BASELINE.json configs: gcm = configs[3] (the headline)")
    ap.add_argument("--models", default="g256", choices=sorted(MODELS), help="network generation: g256/g224 (2018) or the g128/g112 *_BEST set (2021)")
    ap.add_argument("--columns", type=int, default=0, help="total columns over all ranks (0 = the configuration's own count)")
    ap.add_argument("--nlay", type=int, default=0, help="layers (0 = the configuration's own count)")
    ap.add_argument("--cpu-columns", type=int, default=8192, help="columns in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--gather", default="nccl", choices=["p2p", "nccl"], help="N > 1: how the flux slots reach the other ranks (see run_b200)")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="N > 1: cap on the CTAs NCCL may use for the overlapped flux gather (0 = NCCL's default, which measured best: at 8 GPUs "
                         "a cap of 4 / 2 CTAs stretched the gathers from 15 to 41 / 63 ms per pass, profiles/r2p_n8_ctas*.json)")
    ap.add_argument("--gather-pieces", type=int, default=4, help="N > 1: pieces of the rank's shard whose flux gathers overlap the next piece's kernels")
    ap.add_argument("--device-inputs", action="store_true",
                    help="tile the NUNIQUE distinct columns on the device instead of on the host (1e7-column runs: no 40 GB of pinned "
                         "host memory); implies --no-e2e")
    ap.add_argument("--fast-math", dest="fast_math", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--solver-buffer", dest="solver_buffer", type=int, default=0, help="0 auto, 1 shared memory, 2 L2-resident global scratch")
    ap.add_argument("--sw-fast-math", dest="sw_fast_math", type=int, default=1,
                    help="library default: the SW solver without Newton refinements (no measurable accuracy cost, DESIGN.md section 3b)")
    ap.add_argument("--solver-variant", dest="solver_variant", type=int, default=0, help="0 TMA-staged packed solvers, 1 one g-point per lane")
    ap.add_argument("--solver-scratch-mb", dest="solver_scratch_mb", type=int, default=0, help="L2 budget of the packed solvers' reverse-sweep scratch (0 = default)")
    ap.add_argument("--lw-compact-source", dest="lw_compact_source", type=int, default=1,
                    help="1 = LW sources stay factored between gas optics and solver (default), 0 = materialised lay/lev_source")
    ap.add_argument("--solver-warps", dest="solver_warps", type=int, default=0, help="solvers (warps) per CTA in the packed solver kernels (0 = default)")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    args.warmup = max(3, args.warmup) if args.impl == "b200" else max(0, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
