#!/usr/bin/env python
"""Top warp-stall lines of one kernel from an `ncu --set full --import-source on` report.
   ncu -i rep.ncu-rep --page source --csv --print-source sass --kernel-name regex:NAME > src.csv ; python tools/ncu_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
end = next((i for i in range(k + 1, len(rows)) if rows[i] and rows[i][0] in ("Kernel Name", "Address")), len(rows))  # first launch only
hdr, data = rows[k], [r for r in rows[k + 1:end] if len(r) == len(rows[k])]
iS, iI = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
val = lambda r, i: int(r[i]) if r[i].isdigit() else 0
print("kernel:", rows[0][1][:100] if rows[0] else "")
print("samples", sum(val(r, iS) for r in data), "warp instructions", sum(val(r, iI) for r in data), "SASS lines", len(data))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(val(r, hdr.index(h)) for r in data) for h in reasons}
print({a: b for a, b in sorted(agg.items(), key=lambda kv: -kv[1]) if b})
for idx, r in sorted(enumerate(data), key=lambda kr: -val(kr[1], iS))[:n]:
    rs = {h: val(r, hdr.index(h)) for h in reasons if val(r, hdr.index(h))}
    rs = dict(sorted(rs.items(), key=lambda kv: -kv[1])[:2])
    print(str(idx).rjust(5), str(val(r, iS)).rjust(6), str(val(r, iI)).rjust(9), r[1].strip()[:72].ljust(72), rs)
