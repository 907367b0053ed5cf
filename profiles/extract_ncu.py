#!/usr/bin/env python
"""Turn the .ncu-rep captures of a round (gpurun_out/*.ncu-rep, scratch) into the tracked summaries:

    python profiles/extract_ncu.py r1 gpurun_out/r1_measure_full_final.ncu-rep gpurun_out/r1_others_final.ncu-rep

writes profiles/<round>_kernels_ncu.csv (selected `--set full` metrics per captured launch) and
profiles/<round>_traffic.json (DRAM bytes per launch of the fused measurement kernel, which bench.py reports as
roofline.traffic).  Needs `ncu` on PATH (reads reports only; no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def raw_rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    return names, units, rows[hdr + 2:]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def main(argv):
    rnd, reports = argv[1], argv[2:]
    here = os.path.dirname(os.path.abspath(__file__))
    lines = [["report", "kernel", "metric", "unit", "value"]]
    traffic = {}
    for rep in reports:
        names, units, rows = raw_rows(rep)
        kcol = names.index("Kernel Name")
        for r in rows:
            if len(r) <= kcol:
                continue
            kernel = r[kcol]
            got = {}
            for m in METRICS:
                if m in names:
                    c = names.index(m)
                    lines.append([os.path.basename(rep), kernel, m, units[c], r[c]])
                    got[m] = (r[c], units[c])
            if "measure_kernel" in kernel and "dram__bytes_read.sum" in got:
                rd = to_bytes(*got["dram__bytes_read.sum"])
                wr = to_bytes(*got["dram__bytes_write.sum"])
                key = ("measure_kernel<float,f32>" if "LandmarkF" in kernel else
                       "measure_kernel<float>" if "float" in kernel else "measure_kernel<double>")
                traffic[key] = dict(dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes=rd + wr,
                                    grid=got.get("launch__grid_size", ("", ""))[0],
                                    block=got.get("launch__block_size", ("", ""))[0],
                                    source="profiles/%s_kernels_ncu.csv (ncu --set full --clock-control none, one launch, %s)"
                                           % (rnd, os.path.basename(rep)))
    with open(os.path.join(here, "%s_kernels_ncu.csv" % rnd), "w", newline="") as fh:
        csv.writer(fh).writerows(lines)
    if traffic:
        path = os.path.join(here, "%s_traffic.json" % rnd)
        old = {}
        if os.path.exists(path):
            old = json.load(open(path))
        for k, v in traffic.items():
            keep = {kk: vv for kk, vv in old.get(k, {}).items()
                    if kk in ("workload", "particles_per_gpu", "landmarks", "blobs", "dtype")}
            keep.update(v)
            old[k] = keep
        json.dump(old, open(path, "w"), indent=1)
    print("wrote", len(lines) - 1, "metric rows;", "traffic:", {k: v["dram_bytes"] for k, v in traffic.items()})


if __name__ == "__main__":
    main(sys.argv)
