"""DRAM traffic per launch of the four hot-path kernels from an `ncu --set full` report of tools/prof_case.py:
python tools/ncu_traffic.py report.ncu-rep ncol nlay > profiles/<tag>_ncu_traffic.json   (bench.py reads the newest one)"""
import csv, json, subprocess, sys
rep, ncol, nlay = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
def val(r, k):
    return float(r[ix[k]].replace(",", "")) * scale[units[ix[k]]]
names = {"gas_optics_lw": ("gas_optics_tc_kernel<0>", "gas_optics_tc_kernel<0,", "gas_optics_lw"),
         "gas_optics_sw": ("gas_optics_tc_kernel<1>", "gas_optics_tc_kernel<1,", "gas_optics_sw"),
         "lw_solver": ("lw_solver",), "sw_solver": ("sw_solver",)}
res = {}
for r in rows[2:]:
    kn = r[ix["Kernel Name"]]
    for key, subs in names.items():
        if any(s in kn for s in subs):
            res[key] = {"kernel": kn[:60], "ncol_per_launch": ncol, "nlay": nlay, "dram_bytes_read": val(r, "dram__bytes_read.sum"),
                        "dram_bytes_write": val(r, "dram__bytes_write.sum"), "duration_s_under_ncu": val(r, "gpu__time_duration.sum")}
print(json.dumps({"source": "ncu --set full --clock-control none, tools/prof_case.py %d %d (%s)" % (ncol, nlay, rep), "kernels": res}, indent=1))
