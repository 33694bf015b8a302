#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_pair_r1.ncu-rep > profiles/r1_chamfer_pair_full.md

`launches` reads the CSV written by `ncu --metrics gpu__time_duration.sum --csv --log-file …` and prints one row
per kernel (launch count, total and mean device time, share of the captured region).  `full` reads a `--set full`
report through `ncu -i … --page raw --csv` (no GPU needed) and prints the metrics the roofline argument rests on.
"""
import collections
import csv
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    if name.startswith("at::") or "at::native" in name:
        m = re.search(r"(\w+Functor|uniform_kernel|\w+_kernel)", name)
        return "torch:" + (m.group(1) if m else name[:40])
    return re.sub(r"\(.*", "", name)


def launches(path):
    rows = list(csv.DictReader(l for l in open(path, newline="") if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg.setdefault((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"source: `{path}` ({len(rows)} launches, {total / 1e6:.3f} ms captured; ncu per-launch times are cold-cache and "
          "serialised — read the SHARE column)\n")
    print("| kernel | grid | block | launches | total ms | mean us | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (k, g, b), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {g} | {b} | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / total:.1f}% |")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"source: `{path}` (`ncu --set full --clock-control none --import-source on`; read with `--page raw --csv`)\n")
    for r in data:
        name = short(r[col["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        print(f"### launch {r[col['ID']]}: `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in KEY_METRICS:
            if m in col and r[col[m]] != "":
                print(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        print()


def traffic(path, pattern):
    """DRAM bytes (read + written) of ONE launch of every kernel matching `pattern`, summed: the measured traffic of
    one Chamfer forward.  Prints the JSON that bench.py reads (profiles/chamfer_forward_traffic.json)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    seen, kernels, total = set(), [], 0.0
    for r in data:
        name = short(r[col["Kernel Name"]])
        if not re.search(pattern, name) or name in seen:
            continue
        seen.add(name)
        by = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            by += float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
        kernels.append({"kernel": name, "dram_bytes": by, "time_us": r[col["gpu__time_duration.sum"]]})
        total += by
    print(json.dumps({"bytes_per_forward": total, "source": path.replace("gpurun_out/", "") + " (ncu --set full, one launch of each kernel)",
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "full", "traffic"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "chamfer_grid|chamfer_rest")
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
