#!/usr/bin/env python
"""Turns an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the compact per-launch CSV committed
under profiles/, and an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel share table.

  python profiles/summarize.py rep  gpurun_out/prof.ncu-rep        profiles/rNN_name_ncu_full_subset.csv
  python profiles/summarize.py list gpurun_out/launches.csv        profiles/rNN_name_launch_shares.csv [exclude-regex]
(exclude-regex: kernels that are not part of a step — the load-time range calibration, torch's own fill/compare
kernels — are listed with their times but left out of the share column)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "gpc__cycles_elapsed.max.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [d[i] for d in data])
    print(f"{out}: {len(data)} launches x {len(idx)} metrics")


def launch_list(path, out, exclude=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    skip = re.compile(exclude) if exclude else None
    tot = sum(sum(v) for k, v in agg.items() if not (skip and skip.search(k)))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_us", "min_us", "max_us", "share_of_listed_time_pct"])
        for k, v in agg.items():
            share = "excluded" if (skip and skip.search(k)) else f"{100 * sum(v) / tot:.2f}"
            w.writerow([k, len(v), f"{sum(v) / len(v) / 1e3:.2f}", f"{min(v) / 1e3:.2f}", f"{max(v) / 1e3:.2f}", share])
    print(f"{out}: {len(agg)} kernels, {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](*sys.argv[2:])
