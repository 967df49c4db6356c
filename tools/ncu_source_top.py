#!/usr/bin/env python
"""Condenses an `ncu --page source --csv` export: total warp-stall samples by reason and the
instructions that collected the most samples. usage: ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
print(rows[0][1][:120] if len(rows[0]) > 1 else "")
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: 0 for n in stalls}
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    for n in stalls:
        try:
            tot[n] += int(r[col[n]] or 0)
        except ValueError:
            pass
    recs.append((samples, r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                 {n: int(r[col[n]] or 0) for n in stalls if (r[col[n]] or "0") not in ("0", "")}))
all_s = sum(s for s, *_ in recs) or 1
print("instructions:", len(recs), "samples:", all_s, "warp-insts executed:", sum(x[2] for x in recs))
print("stall totals:", {k.replace("stall_", ""): v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
for s, src, ex, st in sorted(recs, key=lambda x: -x[0])[:top]:
    print(f"{100.0 * s / all_s:5.1f}%  {src[:70]:70s} exec={ex} {dict((k.replace('stall_', ''), v) for k, v in st.items())}")
