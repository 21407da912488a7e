"""Top stall sites of a kernel from an .ncu-rep (needs -lineinfo + --import-source on at capture).
usage: python tools/ncu_stalls.py report.ncu-rep [kernel-id (1-based launch index)] [top N]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:100])
hdr = rows[1]
iS, iN, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        body.append(r)
tot = sum(int(r[iS] or 0) for r in body)
print(f"total samples {tot}, {len(body)} SASS instructions")
agg = {}
for i in stall_cols:
    agg[hdr[i]] = sum(int(r[i] or 0) for r in body)
print("stall reasons:", ", ".join(f"{k[6:]}={100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
order = sorted(range(len(body)), key=lambda j: -int(body[j][iS] or 0))[:top]
for j in sorted(order):
    r = body[j]
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{j:5d} {100*int(r[iS])/max(tot,1):5.2f}% exec={r[iN]:>8s}  {r[iSrc].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in reasons if c))
