#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` exports (tools/ncu_export.sh) into small JSON summaries
under profiles/: per launch, the handful of metrics the roofline argument needs.
usage: python tools/ncu_summary.py gpurun_out/add_r01_raw.csv profiles/r01_add.json"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name", "")[:160], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEEP:
            if k in d and d[k] != "":
                try:
                    e[k] = {"value": float(d[k].replace(",", "")), "unit": u.get(k, "")}
                except ValueError:
                    e[k] = {"value": d[k], "unit": u.get(k, "")}
        # warp-stall breakdown: whatever this ncu version calls them, keep those that matter
        for k in hdr:
            if ("issue_stalled" in k and k.endswith(".pct")) or k in (
                    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
                    "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
                    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"):
                try:
                    v = float(d[k].replace(",", ""))
                except (ValueError, KeyError):
                    continue
                if "issue_stalled" not in k or v >= 4.0:
                    e[k] = {"value": v, "unit": u.get(k, "")}
        rd, wr = e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            e["dram_bytes_per_launch"] = rd["value"] * scale.get(rd["unit"], 1.0) + wr["value"] * scale.get(wr["unit"], 1.0)
        out.append(e)
    json.dump({"source": src, "tool": "ncu --set full --clock-control none (one launch, cold cache, serialised)",
               "launches": out}, open(dst, "w"), indent=1)
    for e in out:
        print(e["kernel"][:70], e.get("gpu__time_duration.sum", {}).get("value"), e.get("dram_bytes_per_launch"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
