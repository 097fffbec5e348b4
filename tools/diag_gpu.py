"""GPU diagnostics: where do GPU-vs-oracle differences come from? (run under gpurun)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import helpers as H, oracle as O, nc4min
from rte_rrtmgp_nn_b200 import api, spectral, synth, _lib

ctx = api.default_context(0)

def mlp64(model, x):
    a = x.astype(np.float64)
    n = len(model["W"])
    for l in range(n):
        a = a @ model["W"][l].astype(np.float64) + model["b"][l].astype(np.float64)
        if l < n - 1:
            a = a / (np.abs(a) + 1)
    return a

def stats(name, e):
    e = np.asarray(e, np.float64).ravel()
    print(f"  {name}: max {e.max():.3e} p99.99 {np.percentile(e,99.99):.3e} p99 {np.percentile(e,99):.3e} median {np.median(e):.3e}")

ncol, nlay = 64, 60
atm = synth.make_atmosphere(ncol, nlay, seed=99)
# ---------------- LW tau ----------------
kd = spectral.synthetic_kdist_lw(256)
k_lw = api.ty_gas_optics_rrtmgp(ctx); k_lw.load(kd)
onets, dnets = H.oracle_nets(H.LW_G256), H.device_nets(ctx, H.LW_G256)
ref = O.gas_optics_lw(kd, onets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"])
reff = O.gas_optics_lw(kd, onets, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"], fast=True)
op = api.ty_optical_props_1scl(); op.alloc_1scl(ncol, nlay, k_lw)
src = api.ty_source_func_lw(); src.alloc(ncol, nlay, k_lw)
assert k_lw.gas_optics(atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], H.gas_concs(atm["gases"]), op, src, tlev=atm["tlev"], neural_nets=dnets) == ""
tau = op.tau.cpu().numpy()
m = nc4min.load_nn_model(os.path.join(H.NN_DIR, H.LW_G256[0]))
z64 = mlp64(m, ref["nn_inputs"].reshape(-1, 18))
t64 = ((m["ystd"].astype(np.float64) * z64 + m["ymean"].astype(np.float64)) ** 8 * ref["col_dry"].reshape(-1, 1).astype(np.float64)).reshape(tau.shape)
print("LW tau relative error (plain, no floor):")
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
stats("gpu vs oracle-strict", rel(tau, ref["tau"]))
stats("gpu vs fp64", rel(tau, t64))
stats("oracle-strict vs fp64", rel(ref["tau"], t64))
stats("oracle-fast vs fp64", rel(reff["tau"], t64))
e = rel(tau, ref["tau"]); i = np.unravel_index(np.argmax(e), e.shape)
print("  worst element", i, "tau", tau[i], "sample max", tau[i[0], i[1]].max(), "amp 8*ystd/ymean", 8 * m["ystd"][i[2]] / m["ymean"][i[2]])
w = t64 / t64.max(axis=-1, keepdims=True)
for thr in (1e-2, 1e-4, 1e-6, 1e-8):
    msk = w >= thr
    print(f"  tau >= {thr:g} x sample max: gpu-vs-oracle max {e[msk].max():.3e}  gpu-vs-fp64 max {rel(tau,t64)[msk].max():.3e} oracle-vs-fp64 {rel(ref['tau'],t64)[msk].max():.3e}")

# ---------------- SW ----------------
ks = spectral.synthetic_kdist_sw(224)
k_sw = api.ty_gas_optics_rrtmgp(ctx); k_sw.load(ks)
onets, dnets = H.oracle_nets(H.SW_G224), H.device_nets(ctx, H.SW_G224)
ref = O.gas_optics_sw(ks, onets, atm["play"], atm["plev"], atm["tlay"], atm["gases"])
op = api.ty_optical_props_2str(); op.alloc_2str(ncol, nlay, k_sw)
toa = torch.empty((ncol, 224), device="cuda")
assert k_sw.gas_optics(atm["play"], atm["plev"], atm["tlay"], H.gas_concs(atm["gases"]), op, toa, neural_nets=dnets) == ""
tau = op.tau.cpu().numpy(); ssa = op.ssa.cpu().numpy()
print("SW tau/ssa:")
stats("tau gpu vs oracle", rel(tau, ref["tau"]))
stats("ssa abs diff", np.abs(ssa - ref["ssa"]))
alb = np.repeat(atm["sfc_alb"][:, None], 224, 1)
rup, rdn, rdir = O.rte_sw(True, atm["mu0"], ref["toa_src"], alb, alb, ref["tau"], ref["ssa"], ref["g"])
# (a) oracle solver on GPU tau/ssa: how much of the flux difference is the NN noise?
aup, adn, adir = O.rte_sw(True, atm["mu0"], ref["toa_src"], alb, alb, tau, ssa, ref["g"])
print("SW flux |d| oracle-solver(gpu tau) vs oracle-solver(oracle tau): up %.3e dn %.3e dir %.3e" % (np.abs(aup-rup).max(), np.abs(adn-rdn).max(), np.abs(adir-rdir).max()))
# (b) GPU solver on oracle tau/ssa: how much is the solver?
P = api._ptr
d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(inc=ref["toa_src"], tau=ref["tau"], ssa=ref["ssa"], mu0=atm["mu0"], alb=alb).items()}
mk = lambda: torch.empty((ncol, nlay + 1), device="cuda")
up, dn, dr = mk(), mk(), mk()
for fast in (0, 1):
    ctx.set_flag("fast_math", fast)
    _lib.check(_lib.lib().rrnn_sw_solver_2stream(ctx.h, 224, nlay, ncol, 1, P(d["inc"]), None, P(d["tau"]), P(d["ssa"]), None, P(d["mu0"]), P(d["alb"]), P(d["alb"]), P(up), P(dn), P(dr)))
    print("SW flux |d| gpu-solver(fast=%d) vs oracle-solver, same tau: up %.3e dn %.3e dir %.3e" % (fast, np.abs(up.cpu().numpy()-rup).max(), np.abs(dn.cpu().numpy()-rdn).max(), np.abs(dr.cpu().numpy()-rdir).max()))
ctx.set_flag("fast_math", 0)
# fp64 solve of the same two-stream/adding equations (reference order) for an absolute yardstick
def sw64(tau, ssa, mu0, inc, alb):
    tau = tau.astype(np.float64); w0 = ssa.astype(np.float64); mu0 = mu0.astype(np.float64)[:, None]
    C, L, G = tau.shape
    dirf = np.zeros((C, L + 1, G)); dirf[:, 0] = inc * mu0
    Rdif = np.zeros((C, L, G)); Tdif = np.zeros_like(Rdif); su = np.zeros_like(Rdif); sd = np.zeros_like(Rdif)
    for l in range(L):
        t = tau[:, l]; w = w0[:, l]
        Tn = np.exp(-t / mu0)
        g1 = (8 - w * 5) * .25; g2 = 3 * w * .25; g3 = 0.5 + 0 * w; g4 = 1 - g3
        a1 = g1 * g4 + g2 * g3; a2 = g1 * g3 + g2 * g4
        k = np.sqrt(np.maximum((g1 - g2) * (g1 + g2), 1e-4))
        e = np.exp(-t * k); e2 = e * e; k2e = 2 * k * e
        RT = 1 / (k * (1 + e2) + g1 * (1 - e2))
        Rdif[:, l] = RT * g2 * (1 - e2); Tdif[:, l] = RT * 2 * k * e
        kmu = k * mu0; kmu2 = kmu * kmu; kg3 = k * g3; kg4 = k * g4
        om = 1 - kmu2; dd = np.where(np.abs(om) >= np.finfo(np.float32).eps, om, np.finfo(np.float32).eps)
        RT = w * RT / dd
        Rdir = RT * ((1 - kmu) * (a2 + kg3) - (1 + kmu) * (a2 - kg3) * e2 - k2e * (g3 - a2 * mu0) * Tn)
        Tdir = RT * (k2e * (g4 + a1 * mu0) - Tn * ((1 + kmu) * (a1 + kg4) - (1 - kmu) * (a1 - kg4) * e2))
        Rdir = np.maximum(0, np.minimum(Rdir, 1 - Tn)); Tdir = np.maximum(0, np.minimum(Tdir, 1 - Tn - Rdir))
        su[:, l] = Rdir * dirf[:, l]; sd[:, l] = Tdir * dirf[:, l]; dirf[:, l + 1] = Tn * dirf[:, l]
    albedo = np.zeros((C, L + 1, G)); srcv = np.zeros_like(albedo); den = np.zeros((C, L, G))
    albedo[:, L] = alb; srcv[:, L] = dirf[:, L] * alb
    for l in range(L - 1, -1, -1):
        den[:, l] = 1 / (1 - Rdif[:, l] * albedo[:, l + 1])
        albedo[:, l] = Rdif[:, l] + Tdif[:, l] ** 2 * albedo[:, l + 1] * den[:, l]
        srcv[:, l] = su[:, l] + Tdif[:, l] * den[:, l] * (srcv[:, l + 1] + albedo[:, l + 1] * sd[:, l])
    fdn = np.zeros_like(albedo); fup = np.zeros_like(albedo)
    fup[:, 0] = srcv[:, 0]
    for l in range(1, L + 1):
        fdn[:, l] = (Tdif[:, l - 1] * fdn[:, l - 1] + Rdif[:, l - 1] * srcv[:, l] + sd[:, l - 1]) * den[:, l - 1]
        fup[:, l] = fdn[:, l] * albedo[:, l] + srcv[:, l]
    return fup.sum(-1), (fdn + dirf).sum(-1), dirf.sum(-1)
u64, d64, r64 = sw64(ref["tau"], ref["ssa"], atm["mu0"], ref["toa_src"].astype(np.float64), alb.astype(np.float64))
print("SW vs fp64 (same fp32 tau/ssa): oracle up %.3e dn %.3e | gpu up %.3e dn %.3e" % (np.abs(rup-u64).max(), np.abs(rdn-d64).max(), np.abs(up.cpu().numpy()-u64).max(), np.abs(dn.cpu().numpy()-d64).max()))
