"""Executed warp instructions and stall samples per CUDA SOURCE LINE of one kernel of an ncu report.

    [NCU_KERNEL="<demangled substring>"] python tools/ncu_by_line.py report.ncu-rep <mangled-kernel-substring> <units> [top N] [file.cu]

ncu's CSV source page lists SASS only; nvdisasm -g of the shipped library carries the line table.  Both list the kernel's
instructions in address order, so they are joined by index.  `units` divides the counts (e.g. columns x layers x chunks), so the
table reads "instructions per unit of work".  Inlined code is attributed to the innermost line nvdisasm reports.
"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sub, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
only = sys.argv[5] if len(sys.argv) > 5 else None

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rte_rrtmgp_nn_b200", "lib", "librrnn_b200.so")], cwd=tmp, capture_output=True)
lines = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    sect = [i for i, l in enumerate(dis) if l.lstrip().startswith(".section") and ".text." in l]
    for a, i in enumerate(sect):
        if sub in dis[i]:
            body = dis[i:(sect[a + 1] if a + 1 < len(sect) else len(dis))]
            lines, cur = [], ("?", 0)
            for l in body:
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                elif re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
                    lines.append((cur, l.split("*/", 1)[1].strip().rstrip(";")))
            break
    if lines:
        break
assert lines, "kernel not found in the library"

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blk, cur = None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        want = os.environ.get("NCU_KERNEL")    # substring of the DEMANGLED name in the report (default: derived from the mangled one)
        hit = (want in r[1]) if want else (re.sub(r"[^A-Za-z0-9_]", "", re.split(r"I[LN]", sub)[0][-16:]) in re.sub(r"[^A-Za-z0-9_]", "", r[1]))
        if blk is None and hit:
            blk = cur
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
assert blk, "kernel not found in the report"
ix = {k: i for i, k in enumerate(blk["hdr"])}
assert len(blk["rows"]) == len(lines), (len(blk["rows"]), len(lines))
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_i = tot_s = 0
for (loc, sass), r in zip(lines, blk["rows"]):
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    tot_i += n; tot_s += s
    if only and loc[0] != only:
        loc = (loc[0], 0)
    e = per[loc]
    e[0] += n; e[1] += s
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", sass)
    e[2][m.group(1) if m else "?"] += n
print(f"{blk['name'][:80]}: {tot_i / units:.1f} warp instructions per unit, {tot_s} samples")
src_cache = {}
for loc, (n, s, ops) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    f = os.path.join(ROOT, "rte_rrtmgp_nn_b200", "csrc", loc[0])
    if f not in src_cache:
        src_cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
    text = src_cache[f][loc[1] - 1].strip()[:70] if 0 < loc[1] <= len(src_cache[f]) else ""
    top = " ".join(f"{k}:{v / units:.1f}" for k, v in ops.most_common(4))
    print(f"{loc[0]}:{loc[1]:<5d} {n / units:6.2f}/unit {100.0 * s / max(tot_s, 1):5.1f}% smp | {text:70s} | {top}")
