import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print("value",round(d["value"]),"ms",round(d["ms_per_step"],1), " ".join(f'{k}={v["ms_per_step"]:.1f}ms({v["frac_of_hbm_peak"]:.3f})' for k,v in d["roofline"]["per_kernel"].items()))
