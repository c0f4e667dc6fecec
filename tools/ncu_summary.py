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


STAGE_OF = {"k_zinv_tile": "zinv_tile", "k_yinv_tile": "yinv_tile", "k_yfwd_tile": "yfwd_tile", "k_yfwd_tma": "yfwd_tile",
            "k_zfwd_rk": "zfwd_rk", "k_zfwd_rk_tma": "zfwd_rk", "k_project": "project", "k_project_bulk": "project",
            "k_xpass_gradre": "xpass", "k_xpass_gradre_bulk": "xpass", "k_xpass_cross": "xpass", "k_xpass_cross_bulk": "xpass",
            "k_project_pair": "project", "k_aproject_pair": "project"}


def summarise_rep(tag, reps, workload="hd512"):
    """reps: one or more reports that together cover one substep in launch order (zinv x3, yinv x6, xpass, ...)."""
    import json
    hdr, units, data = [], [], []
    for rep in reps:   # reports may carry different metric columns: re-index every row on the union of the headers
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        for h, u in zip(rows[0], rows[1]):
            if h not in hdr:
                hdr.append(h)
                units.append(u)
        pos = {h: i for i, h in enumerate(rows[0])}
        for r in rows[2:]:
            data.append(({h: r[pos[h]] for h in pos}, {h: rows[1][pos[h]] for h in pos}))   # values, units of THIS report

    class Row(list):
        pass
    width = len(hdr)
    idx = {h: i for i, h in enumerate(hdr)}
    conv = []
    for vals, us in data:
        r = Row([vals.get(h, "") for h in hdr])
        r.units = [us.get(h, "") for h in hdr]
        conv.append(r)
    data = conv
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    out = ["| stage | kernel | launches | time ms | dram read GB | dram write GB | traffic GB/s | dram % of hw peak | LSU data pipe % | fp64 pipe % | issue active % | occupancy % | regs | grid x block | dyn smem KB | top stall reasons (pc samples) |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    agg = collections.OrderedDict()
    n_inv = 0
    for r in data:
        name = r[idx["Kernel Name"]]
        base = name.split("<")[0].split("(")[0].replace("void ", "").replace("sx::", "").strip()
        key = name.split("(")[0].replace("void ", "").replace("sx::", "")
        if base == "k_inv_tma":   # the same kernel serves the z-inverse (first three launches of a substep) and the y-inverse
            stage = "zinv_tile" if n_inv % 9 < 3 else "yinv_tile"
            n_inv += 1 if len(reps) == 1 else 0
            if len(reps) > 1:
                stage = "zinv_tile" if n_inv < 3 else "yinv_tile"
                n_inv += 1
        else:
            stage = STAGE_OF.get(base, base)
        agg.setdefault((stage, key), []).append(r)
    traffic = {}
    for (stage, key), rs in agg.items():
        def avg(metric, conv=None):
            i = idx.get(metric)
            if i is None:
                return float("nan")
            vals = [conv(r[i], r.units[i]) if conv else float(r[i]) for r in rs]
            return sum(vals) / len(vals)
        t = avg("gpu__time_duration.sum", to_ms)
        rd = avg("dram__bytes_read.sum", to_bytes) / 1e9
        wr = avg("dram__bytes_write.sum", to_bytes) / 1e9
        r0 = rs[0]
        tot = sum(float(r0[idx[h]] or 0) for h in stalls) or 1.0
        top = sorted(((float(r0[idx[h]] or 0) / tot, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stalls), reverse=True)[:4]
        smem_i = idx["launch__shared_mem_per_block_dynamic"]
        smem_kb = to_bytes(r0[smem_i], r0.units[smem_i]) / 1e3
        out.append(f"| {stage} | {key} | {len(rs)} | {t:.3f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / (t * 1e-3):.0f} | "
                   f"{avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{avg('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{avg('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{r0[idx['launch__registers_per_thread']]} | {r0[idx['launch__grid_size']]} x {r0[idx['launch__block_size']]} | "
                   f"{smem_kb:.1f} | " + ", ".join(f"{n} {100 * v:.0f}%" for v, n in top) + " |")
        e = traffic.setdefault(stage, {"dram_bytes_per_launch": 0.0, "launches": 0, "kernel": key})
        e["dram_bytes_per_launch"] += (rd + wr) * 1e9 * len(rs)
        e["launches"] += len(rs)
    for e in traffic.values():
        e["dram_bytes_per_launch"] /= e["launches"]
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md")
    with open(path, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, kernels of one RK substep of {workload} ({tag})\n\n"
                "Per-launch averages over the captured launches (cold caches, serialised replays: compare shares and\n"
                "traffic, not absolute times).  `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch.\n"
                "Captured with tools/gpu_ncu3.sh (three reports covering zinv/yinv, xpass/yfwd, zfwd_rk/project).\n\n")
        f.write("\n".join(out) + "\n")
    print("wrote", path)
    # one entry per workload, tied to the kernel sources it was captured from (bench.py drops it when they differ)
    import sys
    sys.path.insert(0, ROOT)
    import bench
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(tpath) as f:
            doc = json.load(f)
        if "workloads" not in doc:
            doc = {"workloads": {}}
    except (OSError, ValueError):
        doc = {"workloads": {}}
    doc["metric"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch"
    doc["workloads"][workload] = {"capture": tag, "sources_hash": bench.sources_hash(), "stages": traffic}
    with open(tpath, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote profiles/ncu_traffic.json")


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
    ap.add_argument("--rep", nargs="+")
    ap.add_argument("--launches")
    ap.add_argument("--workload", default="hd512")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.rep:
        summarise_rep(a.tag, a.rep, a.workload)
    if a.launches:
        summarise_launches(a.tag, a.launches)
