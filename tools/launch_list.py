#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list of
`python bench.py ...` into profiles/: per kernel, launches, total and mean duration, share of the
GPU time of the capture. Per-launch times under ncu are cold-cache and serialised: the SHARE is what
must agree with bench.py's live CUDA-event timings, not the absolute.
usage: python tools/launch_list.py gpurun_out/s12_launches.csv profiles/launches_r01.json"""
import csv
import json
import re
import sys


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Value", "Metric Unit")}
    agg, order = {}, []
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).replace("void ", "").strip()
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        if r[ix["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        key = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
        if key not in agg:
            agg[key] = [0, 0.0]
            order.append(key)
        agg[key][0] += 1
        agg[key][1] += ns
    total = sum(v[1] for v in agg.values())
    out = {"source": src, "tool": "ncu --metrics gpu__time_duration.sum --clock-control none (every launch of the command, serialised)",
           "total_gpu_time_us": round(total / 1e3, 1), "launches": sum(v[0] for v in agg.values()), "kernels": []}
    for key in sorted(order, key=lambda k: -agg[k][1]):
        c, ns = agg[key]
        out["kernels"].append({"kernel": key[0][:200], "grid": key[1], "block": key[2], "launches": c,
                               "total_us": round(ns / 1e3, 1), "mean_us": round(ns / c / 1e3, 2), "share": round(ns / total, 4)})
    json.dump(out, open(dst, "w"), indent=1)
    for k in out["kernels"][:25]:
        print(f'{k["share"]:7.3f} {k["launches"]:5d} x {k["mean_us"]:10.2f} us  {k["kernel"][:110]} {k["grid"]}')


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
