#!/usr/bin/env python
"""Turn gpurun_out/launches.csv (+ optional .ncu-rep) into a committed summary under profiles/.

    python scripts/summarize_ncu.py <tag> [launches.csv] [report.ncu-rep]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        k = row["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    return agg


def main():
    tag = sys.argv[1]
    lpath = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
    rep = sys.argv[3] if len(sys.argv) > 3 else None
    out = [f"# ncu summary {tag}", ""]
    if os.path.exists(lpath):
        agg = launches(lpath)
        tot = sum(v[1] for v in agg.values())
        out += [f"## launch list ({os.path.basename(lpath)}; `ncu --metrics gpu__time_duration.sum --clock-control none`)",
                "", f"total device time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches "
                "(cold-cache, serialised: compare shares, not absolutes)", "",
                "| us | launches | share | kernel |", "|---:|---:|---:|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            mine = " **(ours)**" if "tpspp::" in k else ""
            out.append(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k[:110]}`{mine} |")
        ours = sum(v[1] for k, v in agg.items() if "tpspp::" in k)
        out += ["", f"our kernels: {ours:.1f} us = {100 * ours / tot:.1f}% of device time", ""]
    if rep and os.path.exists(rep):
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        hdr, units = rows[0], rows[1]
        out += [f"## `ncu --set full` ({os.path.basename(rep)})", ""]
        for row in rows[2:]:
            out.append(f"### {row[hdr.index('Kernel Name')]} (launch id {row[hdr.index('ID')]})")
            out.append("")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"- `{w}` = {row[i]} {units[i]}")
            out.append("")
    path = os.path.join(ROOT, "profiles", f"{tag}.md")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
