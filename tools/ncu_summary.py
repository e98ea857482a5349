#!/usr/bin/env python
"""Summarise ncu output for profiles/ (runs in the build container, no GPU needed).

  tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel time shares of the launch list
  tools/ncu_summary.py rep gpurun_out/prof_attn.ncu-rep [...]      -> key metrics of each profiled launch
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void xs::", "").replace("xs::", "").strip()


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    un = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", ""))
        t_us = t / 1e3 if r[un] in ("ns", "nsecond") else (t if r[un] in ("us", "usecond") else t * 1e3)
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += t_us
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms summed device time "
          "(ncu: serialised, cold caches -> compare SHARES)")
    print(f"{'kernel':90s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:90]:90s} {n:5d} {t:11.1f} {t / n:9.1f} {t / total:7.3f}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print(f"## launch id {r[hdr.index('ID')]}: {short(r[hdr.index('Kernel Name')])}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {r[i]:>18s} {units[i]}")
        t = [h for h in hdr if h.startswith("smsp__average_warp") or h.startswith("smsp__average_warps_issue_stalled")]
        stalls = []
        for h in t:
            try:
                stalls.append((float(r[hdr.index(h)].replace(",", "")), h))
            except ValueError:
                pass
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  {h:85s} {v:18.3f}")


if __name__ == "__main__":
    mode, paths = sys.argv[1], sys.argv[2:]
    for p in paths:
        (launches if mode == "launches" else rep)(p)
