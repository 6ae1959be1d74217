"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import csv
import json
import re
import sys

path, command = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[hdr_i + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    name = re.sub(r"\(.*", "", r[iK]).strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
out = {"command": command, "note": "cold-cache, serialised per-launch times: compare SHARES, not absolutes", "total_ms": tot,
       "kernels": [{"kernel": k, "launches": a[0], "ms": round(a[1], 3), "share": round(a[1] / tot, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
