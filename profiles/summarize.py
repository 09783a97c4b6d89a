#!/usr/bin/env python
"""Turns ncu exports brought back from the GPU box (gpurun_out/) into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01c.csv profiles/r01_launches.md
    python profiles/summarize.py full gpurun_out/coarse_r01c.ncu-rep profiles/r01_coarse_full.md [traffic.json [frames per launch]]

`launches`: the `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list of `python bench.py`.
`full`: one `ncu --set full` capture; needs the `ncu` binary (reads the report with --page raw --csv).
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            name = r[4].split("(")[0].split("::")[-1]
            agg.setdefault((name, r[8], r[7]), []).append(float(r[-1]))
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | grid | block | launches | mean us | min us | max us | share of GPU time |\n|---|---|---|---|---|---|---|---|\n")
        for (name, grid, block), v in agg.items():
            f.write("| %s | %s | %s | %d | %.2f | %.2f | %.2f | %.1f %% |\n" % (
                name, grid, block, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, 100 * sum(v) / total))
        f.write("\nSource: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: "
                "compare shares, not absolutes).\n" % src)


def to_bytes(value, unit):
    v = float(value)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(src, dst, traffic_json=None, launch_frames=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("Source: `%s` (ncu --set full --clock-control none --import-source on).\n\n" % src)
        traffic, insts = [], []
        for r in rows[2:]:
            f.write("### %s  grid %s block %s\n\n| metric | unit | value |\n|---|---|---|\n" % (
                r[hdr.index("Kernel Name")].split("(")[0], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, units[i], r[i]))
            f.write("\n")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic.append(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))
            if "smsp__inst_executed.sum" in hdr:
                insts.append(float(r[hdr.index("smsp__inst_executed.sum")].replace(",", "")))
    if traffic_json:
        json.dump({"kernel": rows[2][hdr.index("Kernel Name")].split("(")[0].split("::")[-1],
                   "dram_bytes_per_launch": sum(traffic) / len(traffic),
                   "warp_instructions_per_launch": (sum(insts) / len(insts)) if insts else None,
                   "launches": len(traffic), "launch_frames": int(launch_frames) if launch_frames else 8, "source": src},
                  open(traffic_json, "w"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
