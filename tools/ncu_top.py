"""Summarise the SASS page of an ncu report: hottest instructions by stall samples, with executed counts.
usage: ncu -i rep --page source --csv | python tools/ncu_top.py [kernel-substring] [N]"""
import csv, sys
sub = sys.argv[1] if len(sys.argv) > 1 else ""
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(sys.stdin))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks:
    if sub not in b["name"]:
        continue
    h = b["hdr"]; ix = {k: i for i, k in enumerate(h)}
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in b["rows"]); tin = sum(int(r[ix["Instructions Executed"]]) for r in b["rows"])
    print("==", b["name"][:90], "| SASS instrs", len(b["rows"]), "| samples", tot, "| warp-instr executed", tin)
    agg = {k: sum(int(r[ix[k]]) for r in b["rows"]) for k in stall_cols}
    print("   stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][ix["# Samples"]]))[:N]
    for i in sorted(order):
        r = b["rows"][i]
        st = sorted(((int(r[ix[k]]), k) for k in stall_cols), reverse=True)[:2]
        print(f"  #{i:5d} smp {int(r[ix['# Samples']]):6d} exec {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
