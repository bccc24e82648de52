#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python profiles/summarize.py full gpurun_out/prof_scan_r1.ncu-rep profiles/r1_scan_topk_full.md
"""
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
       "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
       "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
       "dram__bytes_read.sum.per_second", "launch__cluster_size",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, agg, order = None, {}, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        v = float(d["Metric Value"].replace(",", ""))
        if name not in agg:
            agg[name] = []
            order.append(name)
        agg[name].append(v)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none), source `{src}`\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:110]}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/total:.1f}% |\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture, source `{src}` (not tracked; regenerate with the command in profiles/README.md)\n\n")
        for r in rows[2:]:
            if r[idx["gpu__time_duration.sum"]] in ("", "-nan", "nan"):
                continue
            f.write(f"## {r[idx['Kernel Name']]}  (grid {r[idx['launch__grid_size']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEY:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
