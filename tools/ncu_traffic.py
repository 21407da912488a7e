"""Turn an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv) of one denoise
step into profiles/r1_gemm_traffic.json: DRAM bytes per launch of the dominant kernel (f8_gemm_kernel), plus the
per-kernel table.   usage: python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r1_gemm_traffic.json"""
import csv
import json
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
per = defaultdict(lambda: defaultdict(float))
launches = defaultdict(set)
for r in rows:
    name = r["Kernel Name"].split("(")[0]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per[name][r["Metric Name"]] += val * mult
    launches[name].add(r["ID"])
out = {"kernels": {}}
for name, m in per.items():
    n = len(launches[name])
    out["kernels"][name] = {"launches": n, "dram_bytes_read": m.get("dram__bytes_read.sum", 0.0),
                            "dram_bytes_write": m.get("dram__bytes_write.sum", 0.0),
                            "time_ns": m.get("gpu__time_duration.sum", 0.0)}
g = [v for k, v in out["kernels"].items() if "f8_gemm_kernel" in k]
tot = sum(v["dram_bytes_read"] + v["dram_bytes_write"] for v in g)
n = sum(v["launches"] for v in g)
out["dram_bytes_per_launch"] = tot / max(n, 1)
out["gemm_launches"] = n
out["gemm_dram_bytes_per_step"] = tot
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_launch", "gemm_launches", "gemm_dram_bytes_per_step")}))
