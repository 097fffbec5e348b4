#!/bin/bash
# Column-count sweep (BASELINE configs[4]): LW+SW at 91 layers, 1e3 ... 1e7 columns on one GPU (the 1e7 point tiles its inputs on the device and has no e2e leg) -> profiles/<tag>_sweep_L91.jsonl
tag=${1:-r1}
out=gpurun_out/${tag}_sweep_L91.jsonl
: > $out
for n in 1000 3000 10000 30000 100000 300000 1000000 3000000; do
  timeout 300 python bench.py --columns $n --nlay 91 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' >> $out
done
timeout 400 python bench.py --columns 10000000 --nlay 91 --steps 3 --warmup 3 --no-cpu-baseline --device-inputs 2>/dev/null | grep '^{' >> $out
python - <<'PY' $out
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l); k = d["roofline"]["per_kernel"]
    print(f'{d["config"]["ncol_total"]:>8d} cols: {d["value"]:>10.0f} col/s device, {(d["e2e"] or {"value": float("nan")})["value"]:>10.0f} col/s e2e | ' +
          " ".join(f'{n}={v["frac_of_hbm_peak"]:.2f}' for n, v in k.items()))
PY
