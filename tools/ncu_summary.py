#!/usr/bin/env python
"""Summarise an ncu report (--set full) and/or an ncu launch list into profiles/<tag>_*.md|csv.

    python tools/ncu_summary.py <tag> [--rep gpurun_out/<tag>_prof.ncu-rep] [--launches gpurun_out/<tag>_launches.csv]
"""
import argparse
import collections
import csv
import io
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_hw_peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def to_bytes(v, unit):
    f = float(v)
    u = unit.lower()
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_ms(v, unit):
    f = float(v)
    return f * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit.lower(), 1.0)


def summarise_rep(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | launches | time ms | dram read GB | dram write GB | traffic GB/s | dram % of hw peak | fp64 pipe % | issue active % | occupancy % | regs | grid x block | dyn smem KB |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    agg = collections.OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]]
        key = name.split("(")[0].replace("void ", "").replace("sx::", "")
        agg.setdefault(key, []).append(r)
    for key, rs in agg.items():
        def avg(metric, conv=None):
            i = idx.get(metric)
            if i is None:
                return float("nan")
            vals = [conv(r[i], units[i]) if conv else float(r[i]) for r in rs]
            return sum(vals) / len(vals)
        t = avg("gpu__time_duration.sum", to_ms)
        rd = avg("dram__bytes_read.sum", to_bytes) / 1e9
        wr = avg("dram__bytes_write.sum", to_bytes) / 1e9
        r0 = rs[0]
        out.append(f"| {key} | {len(rs)} | {t:.3f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / (t * 1e-3):.0f} | "
                   f"{avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{avg('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{r0[idx['launch__registers_per_thread']]} | {r0[idx['launch__grid_size']]} x {r0[idx['launch__block_size']]} | "
                   f"{float(r0[idx['launch__shared_mem_per_block_dynamic']]) if units[idx['launch__shared_mem_per_block_dynamic']].lower() == 'kbyte' else to_bytes(r0[idx['launch__shared_mem_per_block_dynamic']], units[idx['launch__shared_mem_per_block_dynamic']]) / 1e3:.1f} |")
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md")
    with open(path, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one RK substep of HD 512^3 ({tag})\n\n"
                "Per-launch averages over the captured launches (cold caches, serialised replays: compare shares and\n"
                "traffic, not absolute times).  `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch.\n\n")
        f.write("\n".join(out) + "\n")
    print("wrote", path)


def summarise_launches(tag, path_csv):
    lines = [l for l in open(path_csv) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("sx::", "")
        t = to_ms(r[idx["Metric Value"]].replace(",", ""), r[idx["Metric Unit"]])
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in tot.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {t:.3f} | {100 * t / total:.1f}% |")
    path = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
    with open(path, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps 1 --warmup 3` ({tag})\n\n"
                "First 600 launches (set-up of the synthetic state + warm-up + timed step).  Serialised, cold-cache times: the SHARES are\nwhat is comparable with the CUDA-event stage times in the bench JSON.\n\n")
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.rep:
        summarise_rep(a.tag, a.rep)
    if a.launches:
        summarise_launches(a.tag, a.launches)
