#!/usr/bin/env python
"""Static SASS evidence of the built library: per-kernel instruction-mnemonic counts and resource usage.

    python tools/sass_evidence.py [pattern ...] > profiles/<round>_sass_evidence.md

Runs `cuobjdump -sass` and `cuobjdump -res-usage` on specter_b200/csrc/libspecter_b200.so (no GPU needed), demangles the
kernel names and prints one row per kernel whose name contains one of the patterns (default: the kernels the bench
launches at 512^3).  UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = cp.async.bulk.tensor loads / stores, SYNCS = mbarrier
operations, LDGSTS = cp.async, HMMA / DMMA / UTCMMA = tensor-core instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "specter_b200", "csrc", "libspecter_b200.so")
DEFAULT = ["k_inv_tma<512, 4, 3, 1>", "k_inv_tma<512, 4, 2, 2>", "k_xpass_gradre_bulk<512, 3, 3, 4>", "k_xpass_gradre_bulk<512, 4, 3, 4>",
           "k_xpass_cross_bulk<512, 3, 4>", "k_yfwd_tile<512, 4, 3, 1>", "k_zfwd_rk<512, 4, 2, 1, 1>",
           "k_project_pair<512, 4>", "k_aproject_pair<512", "k_mhd_curls", "k_zgemm"]
COLS = ["UBLKCP", "UTMALDG", "UTMASTG", "UBLKPF", "SYNCS", "LDGSTS", "LDS", "STS", "SHFL", "DFMA", "DADD", "DMUL", "BAR"]
TENSOR = ("HMMA", "DMMA", "IMMA", "UTCMMA", "UTCHMMA", "QMMA")


def demangle(names):
    out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)      # cu++filt prints template values as (int)512
    name = re.sub(r"\(.*$", "", name)
    return name.replace("sx::", "").replace("(anonymous namespace)::", "")


def main():
    pats = sys.argv[1:] or DEFAULT
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
    names = demangle(list(counts))
    tensor_total = sum(c[t] for c in counts.values() for t in TENSOR)
    print(f"# SASS evidence (`cuobjdump -sass` / `-res-usage` of specter_b200/csrc/libspecter_b200.so, sm_100a)\n")
    print(f"{len(counts)} kernels in the library; tensor-core instructions ({', '.join(TENSOR)}) in the whole library: {tensor_total} "
          f"(the path is FP64 transforms and elementwise work: nothing is a dense contraction).\n")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("Whole library: " + ", ".join(f"{k} {tot[k]}" for k in COLS) + "\n")
    print("| kernel | instr | regs | static smem (1 KB reserved) | local (spill) B | " + " | ".join(COLS) + " |")
    print("|---|---|---|---|---|" + "---|" * len(COLS))
    for pat in pats:
        for mangled, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
            nm = short(names[mangled])
            if pat in nm:
                reg, sh, loc = usage.get(mangled, (-1, -1, -1))
                print(f"| `{nm}` | {c['_total']} | {reg} | {sh} | {loc} | " + " | ".join(str(c[k]) for k in COLS) + " |")


if __name__ == "__main__":
    main()
