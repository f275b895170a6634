"""Turns one evidence pass of tools/final_profile.sh (files gpurun_out/<tag>_*) into the tracked summaries under profiles/:
python tools/collect_profiles.py <tag>   (run in the build container; needs ncu for the .ncu-rep export)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for name in ("bench_n1.json", "bench_ref.json", "clocks.csv", "pytest.log", "launches_traffic_n26.csv", "scale_8.json", "scale_4.json"):
    p = os.path.join(src, f"{tag}_{name}")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, f"{tag}_{name}"))

# launch list -> per-kernel time share and DRAM bytes per launch
launch_csv = os.path.join(src, f"{tag}_launches_traffic_n26.csv")
if os.path.exists(launch_csv):
    text = open(launch_csv).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = {}
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, {"launches": set(), "ns": 0.0, "bytes": 0.0})
        a["launches"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            a["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        else:
            a["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    total = sum(a["ns"] / len(a["launches"]) for a in agg.values())
    out = [{"kernel": k, "launches": len(a["launches"]), "ms_per_launch": a["ns"] / len(a["launches"]) / 1e6,
            "share_of_step": a["ns"] / len(a["launches"]) / total, "dram_bytes_per_launch": a["bytes"] / len(a["launches"])}
           for k, a in agg.items()]
    json.dump(out, open(os.path.join(dst, f"{tag}_launch_shares_traffic_n26.json"), "w"), indent=1)
    traffic = {"_source": f"profiles/{tag}_launches_traffic_n26.csv: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
               "dram__bytes_write.sum --clock-control none -k regex:loss_tma_kernel|fp_kernel|gram64 -c 16 on `python bench.py "
               "--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs` (N=2^26, d=64); dram read+write bytes per launch"}
    for o in out:
        for key, pat in (("loss_kernel", "loss_tma_kernel"), ("fp_kernel_f64", "fp_kernel_f64"), ("gram64_kernel", "gram64_kernel")):
            if pat in o["kernel"]:
                traffic[key] = int(o["dram_bytes_per_launch"])
    json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    for o in out:
        print(o)

# ncu --set full capture -> one row per metric, one column per kernel
rep = os.path.join(src, f"{tag}_prof_n23.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    keep = [i for i, h in enumerate(hdr) if "__" in h]
    with open(os.path.join(dst, f"{tag}_ncu_full_summary_n23.csv"), "w") as f:
        f.write("metric,unit," + ",".join('"' + r[ki].split("(")[0][-40:] + '"' for r in data) + "\n")
        for i in keep:
            f.write(f"{hdr[i]},{units[i]}," + ",".join(r[i].replace(",", "") for r in data) + "\n")
    print("ncu summary:", len(keep), "metrics x", len(data), "kernels")
