#!/usr/bin/env python3
"""Copy the judged files of a tools/gpu_r2final.sh run from gpurun_out/<tag>/ into profiles/ (round-2 names) and derive
profiles/r02_ncu_traffic.json (DRAM bytes of one k_forces launch per workload) from the full captures.
usage: tools/collect_final.py <tag>"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2final"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
for f in sorted(os.listdir(src)):
    p = os.path.join(src, f)
    if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(p) > 0:
        shutil.copy(p, os.path.join(dst, "r02_" + f))
    elif f.startswith("launches_") and f.endswith(".csv"):
        shutil.copy(p, os.path.join(dst, "r02_" + f))
traffic = {}
for w in ("impact", "sedov", "giant_hydro"):
    rep = os.path.join(src, f"full_{w}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = os.path.join(dst, f"r02_ncu_full_{w}_summary.csv")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, out])
    rows = list(csv.reader(open(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if "k_forces" in r[0]:
            def val(name):
                i = hdr.index(name)
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
                return float(r[i]) * scale
            traffic[w] = {"k_forces_dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                          "k_forces_ms_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])
                          * {"ms": 1.0, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
                          "source": f"profiles/r02_ncu_full_{w}_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, "
                                    "warm evaluation, step-0 state in cell order)"}
if traffic:
    json.dump(traffic, open(os.path.join(dst, "r02_ncu_traffic.json"), "w"), indent=1)
log = os.path.join(src, "pytest_gpu.log")
if os.path.exists(log):
    lines = open(log).read().splitlines()
    with open(os.path.join(dst, "r02_gputest_summary.txt"), "w") as fh:
        fh.write(f"# python -m pytest tests -m gpu -q on one B200 (tools/gpu_r2final.sh {tag}); tail of the log\n")
        fh.write("\n".join(lines[-25:]) + "\n")
print(json.dumps(traffic, indent=1))
