"""Build the CUDA extension in-tree (and, for tests only, the CPU emulation of the kernels).

    python -m specter_b200.build          # nvcc -> specter_b200/csrc/libspecter_b200.so (sm_100a)
    python -m specter_b200.build --emu    # g++  -> tests/emu/_build/libspecter_emu.so (tests only)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["sx_api.cu", "sx_kernels_fft.cu", "sx_kernels_ops.cu", "sx_rkstep.cu", "sx_solvers.cu", "sx_fused.cu", "sx_fused_tiles.cu",
           "sx_fused_zfwd.cu", "sx_fused_x.cu", "sx_fused_project.cu", "sx_tma.cu", "sx_io.cu", "sx_comm.cu", "sx_walls_diag.cu", "sx_boots.cu"]
LIB = os.path.join(CSRC, "libspecter_b200.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu", "_build")
EMU_LIB = os.path.join(EMU_DIR, "libspecter_emu.so")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps():
    out = _sources()
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    out.append(os.path.join(ROOT, "include", "specter_b200.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB, _deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-Xlinker", "-Bsymbolic", "-o", LIB] + objs + ["-lcudart", "-lnccl", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


def build_emu(force: bool = False) -> str:
    """CPU-thread emulation of the same kernel sources; test infrastructure only."""
    shim = os.path.join(ROOT, "tests", "emu", "cuda_emu.h")
    if not force and not _stale(EMU_LIB, _deps() + [shim]):
        return EMU_LIB
    os.makedirs(EMU_DIR, exist_ok=True)
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(EMU_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        # _FORTIFY_SOURCE off: the fibers switch stacks with _longjmp, which the fortified longjmp check refuses
        cmd = ["g++", "-O2", "-std=c++20", "-fPIC", "-DSX_EMU", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-include", shim,
               "-x", "c++", "-c", src, "-o", obj, "-Wno-unknown-pragmas"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode:
            sys.stderr.write(out)
            raise RuntimeError("g++ (emu) failed: " + " ".join(cmd))
    subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", EMU_LIB] + objs + ["-lpthread", "-latomic"])
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
