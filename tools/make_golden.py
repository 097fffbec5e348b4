"""Generate tests/golden/oracle_golden.npz: outputs of the strict fp32 oracle on fixed seeded inputs.
These pin the ORACLE against regressions (the reference itself holds no golden vectors for this path -- SURVEY.md
section 4 -- and cannot be run here); GPU parity is then checked against the oracle.  Re-run only deliberately."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import helpers as H, oracle as O
from rte_rrtmgp_nn_b200 import spectral, synth, rfmip

out = {}
kd, ks = spectral.synthetic_kdist_lw(256), spectral.synthetic_kdist_sw(224)
lw, sw = H.oracle_nets(H.LW_G256), H.oracle_nets(H.SW_G224)

def run(tag, atm, tsi=None):
    ncol = atm["play"].shape[0]
    go = O.gas_optics_lw(kd, lw, atm["play"], atm["plev"], atm["tlay"], atm["tsfc"], atm["gases"], tlev=atm["tlev"])
    emis = np.repeat(atm["sfc_emis"][:, None], 16, 1)
    up, dn = O.rte_lw(kd, atm["top_at_1"], go["tau"], go["lay_source"], go["lev_source"], go["sfc_source"], emis)
    gs = O.gas_optics_sw(ks, sw, atm["play"], atm["plev"], atm["tlay"], atm["gases"])
    toa = gs["toa_src"]
    if tsi is not None:
        d = np.float32(0)
        for v in ks["solar_source"]:
            d = np.float32(d + v)
        toa = (toa * tsi[:, None] / d).astype(np.float32)
    alb = np.repeat(atm["sfc_alb"][:, None], 224, 1)
    su, sd, sr = O.rte_sw(atm["top_at_1"], atm["mu0"], toa, alb, alb, gs["tau"], gs["ssa"], gs["g"])
    out.update({f"{tag}_lw_up": up, f"{tag}_lw_dn": dn, f"{tag}_sw_up": su, f"{tag}_sw_dn": sd, f"{tag}_sw_dir": sr,
                f"{tag}_lw_tau_sum": go["tau"].sum(axis=-1, dtype=np.float64), f"{tag}_sw_tau_sum": gs["tau"].sum(axis=-1, dtype=np.float64),
                f"{tag}_lw_tau_slice": go["tau"][:2, ::7, ::5].copy(), f"{tag}_sw_ssa_slice": gs["ssa"][:2, ::7, ::5].copy(),
                f"{tag}_lay_source_slice": go["lay_source"][:2, ::7, ::5].copy()})

run("tiny", synth.make_atmosphere(4, 5, seed=1))
run("synth60", synth.make_atmosphere(6, 60, seed=2))
cols = np.arange(0, 1800, 10)
atm = rfmip.load(columns=cols)
out["rfmip_columns"] = cols
run("rfmip", atm, tsi=atm["tsi"])
dst = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst) / 1e3, "KB")
print("RFMIP mean LW flux dn", out["rfmip_lw_dn"].mean(), "up", out["rfmip_lw_up"].mean(), "SW dn", out["rfmip_sw_dn"].mean(), "up", out["rfmip_sw_up"].mean())
