"""Static SASS evidence: per-kernel counts of the Blackwell-specific opcodes in the shipped library.

    python tools/sass_opcodes.py [out.txt]        (default: print)

UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store,
UBLKCP = cp.async.bulk (1-D), SYNCS = mbarrier ops, FFMA2 / FMUL2 / FADD2 = packed fp32x2 arithmetic, CCTL = discard.L2.
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rte_rrtmgp_nn_b200", "lib", "librrnn_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "CCTL", "FFMA", "HMMA", "total"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0]] += 1
        cur["total"] += 1
names = demangle(list(per))
short = lambda n: re.sub(r"\(.*", "", n).replace("rrnn::", "")
rows = [("kernel", *KEYS)]
tot = collections.Counter()
for (mangled, c), n in zip(per.items(), names):
    tot.update(c)
    if any(c[k] for k in KEYS[:7]) or c["FFMA2"] + c["FMUL2"] > 50:
        rows.append((short(n)[:64], *[str(c[k]) for k in KEYS]))
rows.append((f"ALL {len(per)} kernels of librrnn_b200.so", *[str(tot[k]) for k in KEYS]))
w = [max(len(r[i]) for r in rows) for i in range(len(rows[0]))]
text = "\n".join("  ".join(c.ljust(w[i]) if i == 0 else c.rjust(w[i]) for i, c in enumerate(r)) for r in rows) + "\n"
text = f"# cuobjdump -sass rte_rrtmgp_nn_b200/lib/librrnn_b200.so, instruction counts per kernel (tools/sass_opcodes.py)\n" + text
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
else:
    sys.stdout.write(text)
