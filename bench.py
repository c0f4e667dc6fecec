#!/usr/bin/env python
"""Benchmark of the per-RK-substep hot path (BASELINE.json metric: grid-point RK substeps per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload hd512|bouss512|mhd512|hd2048|...]

A "step" is one full Runge-Kutta time step (rkstep1 + `ord` substeps, specter.fpp:1142-1161) of the solver on
synthetic initial conditions; value = nx*ny*nz*ord*K / device seconds (continuation planes included, as the
reference's benchmark.txt counts them).  Default workload: BASELINE.json configs[1] (HD channel flow 512^3, FP64,
RK4) on one B200 and 512^3 points per GPU on 2 and 4 (weak scaling of that configuration); configs[4] (HD 2048x2048x1024) on 8
GPUs.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
TABLES = os.path.join(ROOT, "tests", "golden", "tables")

WORKLOADS = {
    # name: (nx, ny, nz, ord, dt, description)
    "hd512": (512, 512, 512, 4, 2e-4, "HD channel flow 512x512x512 FP64 RK4, no-slip walls, FC-Gram C=25 d=5 (BASELINE configs[1])"),
    "hd256": (256, 256, 256, 4, 5e-4, "HD channel flow 256^3 FP64 RK4 (reduced; not the headline config)"),
    "hd64": (64, 64, 64, 2, 1e-3, "HD 64^3 RK2 (BASELINE configs[0], parity config)"),
    "hd1024": (1024, 1024, 512, 4, 1e-4, "HD 1024x1024x512 FP64 RK4 (needs >= 2 GPUs)"),
    "hd1024c": (1024, 1024, 1024, 4, 1e-4, "HD 1024x1024x1024 FP64 RK4 (the per-GPU problem of BASELINE configs[4] on 2 GPUs)"),
    "hd2048h": (2048, 1024, 1024, 4, 5e-5, "HD 2048x1024x1024 FP64 RK4 (the per-GPU problem of BASELINE configs[4] on 4 GPUs)"),
    "hd2048": (2048, 2048, 1024, 4, 5e-5, "HD 2048x2048x1024 FP64 RK4 (BASELINE configs[4]; 8 GPUs)"),
    "hdxy2048": (2048, 2048, 64, 4, 5e-5, "HD 2048x2048x64 (single-GPU tuning grid for the length-2048 x / y kernels of configs[4]; not a BASELINE config)"),
    "hdz1024": (256, 256, 1024, 4, 5e-5, "HD 256x256x1024 (single-GPU tuning grid for the length-1024 z kernels of configs[4]; not a BASELINE config)"),
    "bouss512": (512, 512, 512, 4, 2e-4, "BOUSS Rayleigh-Benard 512x512x512 FP64 RK4, no-slip + constant-temperature walls"),
    "bouss1024": (1024, 1024, 512, 4, 1e-4, "BOUSS Rayleigh-Benard 1024x1024x512 FP64 RK4 (BASELINE configs[2]; 8 GPUs)"),
    "mhd512": (512, 512, 512, 4, 2e-4, "MHD vector potential 512x512x512 FP64 RK4, no-slip + conducting walls (BASELINE configs[3])"),
}
# --gpus N without --workload: the headline configuration (BASELINE configs[1]) on one GPU, its weak-scaling series of
# 512^3 points per GPU on 2 and 4 (1024x512x512, 1024x1024x512), and the north star's scaling configuration
# (BASELINE configs[4], 2048x2048x1024 = 4 x 512^3 points per GPU) on 8.  The per-GPU problem of configs[4] on fewer
# GPUs is --workload hd1024c (2) / hd2048h (4); profiles/r2h, r2i4.
DEFAULT_WORKLOAD = {1: "hd512", 2: "hd512", 4: "hd512", 8: "hd2048"}
CZ, OZ, NU = 25, 5, 1e-3
KAPPA, MU = 1e-3, 5e-3
# algorithmic HBM bytes per grid-point-substep (55 F / 73 F / 98 F, F = 8 B/pt)
B_ALG_BY_SOLVER = {"hd": 440.0, "bouss": 584.0, "mhd": 784.0}
B_ALG = 440.0


def stage_bytes_per_pt(r, solver="hd"):
    """Algorithmic bytes per grid point and SUBSTEP of each kernel family as built (DESIGN.md "Kernels"): F = 8 B per
    point per full-field read or write; r = physical rows / nz (only physical rows cross the transposition).
"""
    F = 8.0
    if solver == "hd":
        return {"zinv_tile": (3 + 6 * r) * F, "yinv_tile": 15 * r * F, "xpass": 12 * r * F, "yfwd_tile": 6 * r * F,
                "zfwd_rk": (12 + 3 * r) * F, "project": 7 * F}
    if solver == "bouss":   # theta rides with v: 4 components, its z-forward reads v_z and v_z's reads theta
        return {"zinv_tile": (4 + 8 * r) * F, "yinv_tile": 20 * r * F, "xpass": 16 * r * F, "yfwd_tile": 8 * r * F,
                "zfwd_rk": (18 + 4 * r) * F, "project": 7 * F}
    # MHD: one pass for the nine curls (6 F in, 9 F out), 12 plain inverse fields (v, omega, B, J), two cross-product x
    # passes (12 + 6 lines in, 3 + 3 out), two pencil kernels at the end (velocity: 3 in, 3 + p' out; potential: 3 in, 3 + ph out)
    return {"elementwise": 15 * F, "zinv_tile": 12 * (1 + r) * F, "yinv_tile": 24 * r * F, "xpass": 24 * r * F, "yfwd_tile": 12 * r * F,
            "zfwd_rk": (24 + 6 * r) * F, "project": 14 * F}


# the files that define the fused-substep kernels (and the headers they include): what a kernel's DRAM traffic depends on
KERNEL_SOURCES = ("sx_common.cuh", "sx_fft.cuh", "sx_tma.cuh", "sx_plan.h", "sx_fused.h", "sx_fused_tiles.cu", "sx_fused_x.cu",
                  "sx_fused_zfwd.cu", "sx_fused_project.cu", "sx_solvers.cu")


def gpu_view(text):
    """A source file as nvcc sees it with respect to SX_EMU: the regions that only the CPU-thread emulation of the tests
    compiles (`#ifdef SX_EMU` ... / the `#else` side of `#ifndef SX_EMU`) are dropped together with the directives that
    delimit them, so that an edit of the emulation side does not look like a change of the kernels."""
    out, stack = [], []          # stack entries: None for an unrelated #if, else [keep_this_branch]
    for line in text.splitlines():
        t = line.strip()
        if t.startswith("#if"):
            toks = t.split()
            if len(toks) >= 2 and toks[1] == "SX_EMU" and toks[0] in ("#ifdef", "#ifndef"):
                stack.append([toks[0] == "#ifndef"])
                continue
            stack.append(None)
        elif t.startswith("#else") and stack and stack[-1] is not None:
            stack[-1][0] = not stack[-1][0]
            continue
        elif t.startswith("#endif") and stack:
            if stack.pop() is not None:
                continue
        if all(f is None or f[0] for f in stack):
            out.append(line)
    return "\n".join(out) + "\n"


def sources_hash(root=None):
    """Content hash of the kernel sources as the GPU build sees them (gpu_view): ties a committed ncu capture
    (profiles/ncu_traffic.json) to the code it was taken from (the GPU box has no .git).  Host-side orchestration
    (sx_fused.cu, sx_comm.cu, sx_api.cu ...) is not part of it: it does not change what a launch of a kernel reads and
    writes."""
    h = hashlib.sha1()
    d = os.path.join(root or ROOT, "specter_b200", "csrc")
    for f in KERNEL_SOURCES:
        h.update(f.encode())
        with open(os.path.join(d, f), "r") as fh:
            h.update(gpu_view(fh.read()).encode())
    return h.hexdigest()[:12]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler is started before the warm-up
    (nvidia-smi needs a few hundred ms to deliver its first line, longer than a short timed region) and its lines carry
    timestamps; the report uses the lines that fall inside the timed region, or, when the region was too short to catch
    one, the lines since the start of the warm-up -- the same kernels under the same load -- and says which."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/sx_clocks_{os.getpid()}.csv"
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)    # let the line that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows = []
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), c[4:8]))
            except ValueError:
                continue
        try:
            os.remove(self.path)
        except OSError:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or r[0]) + 0.05]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (the timed region was shorter than the sampling latency)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in inside),
                "reasons": sorted(reasons), "samples": len(inside), "window": window}


def _fill_low_modes(plan, buf, seed, c, kmax=4):
    """Random low-wavenumber modes with a wall-vanishing z envelope in the mixed (z,ky,kx) domain, written into the
    (zeroed) host buffer `buf`; returns the index tuples that were written so that the caller can clear them again."""
    import numpy as np
    nxl, ny, nz = plan.cshape
    nph = nz - plan.Cz
    rng = np.random.default_rng(seed + 7919 * c + plan.ista)
    z = np.arange(nph) / (nph - 1.0)
    env = np.sin(np.pi * z) ** 2
    N = float(plan.nx) * plan.ny * plan.nz
    touched = []
    for i in range(nxl):
        kx = plan.ista - 1 + i
        if kx > kmax:
            break
        for j in list(range(0, kmax + 1)) + list(range(ny - kmax, ny)):
            if kx == 0 and j > ny // 2:
                continue
            amp = (rng.standard_normal() + 1j * rng.standard_normal()) * (N / plan.nz) * 0.05
            prof = env * np.cos(np.pi * (1 + (i + j + c) % 3) * z)
            buf[i, j, :nph] = amp * prof
            touched.append((i, j))
            if kx == 0:
                if j == 0:
                    buf[i, j, :nph] = amp.real * prof
                else:
                    buf[i, ny - j, :nph] = np.conj(buf[i, j, :nph])
                    touched.append((i, ny - j))
    return touched


class HostScratch:
    """ONE reusable host field per rank for the set-up (the 2048x2048x1024 fields are 4.3 GB per rank each)."""

    def __init__(self, plan):
        import numpy as np
        self.buf = np.zeros(plan.cshape, dtype=np.complex128)

    def upload(self, plan, dev, seed, c):
        touched = _fill_low_modes(plan, self.buf, seed, c)
        dev.put(self.buf)
        for (i, j) in touched:
            self.buf[i, j, :] = 0.0


def device_state(plan, solver):
    """Synthetic initial condition built IN PLACE in the plan-owned device state with the product API only (this is
    not the initialv.f90 recipe of SURVEY 8(d): values do not affect timing -- no data-dependent control flow -- and
    the parity legs use the oracle's recipe): low modes -> continued z-FFT -> the solver's own boundary-condition /
    projection operators; uniform body force f0 = 1 in x (initialfv.f90:25-31).  Host side: one field per rank."""
    hs = HostScratch(plan)
    N = float(plan.nx) * plan.ny * plan.nz
    if solver == "hd":
        fld, put = plan.hd_field, plan.hd_put_state
        v, pr, fx = [fld(i) for i in range(3)], fld(3), fld(4)
    elif solver == "bouss":
        fld = plan.bouss_field
        v, pr, fx = [fld(i) for i in range(3)], fld(3), fld(4)
    else:
        fld = plan.mhd_field
        v, pr, fx = [fld(i) for i in range(3)], fld(3), fld(4)
    for c in range(3):
        hs.upload(plan, v[c], 1234, c)
        plan.fftp1d_real_to_complex_z(v[c])
    plan.v_imposebc_and_project(v[0], v[1], v[2], pr, plan.ord)
    pr.put(hs.buf)                       # p' = 0 like the reference's start (buf is all zero here)
    if plan.ista == 1:
        hs.buf[0, 0, 0] = 1.0 * N
    fx.put(hs.buf)
    hs.buf[0, 0, 0] = 0.0
    if solver == "bouss":
        th = fld(10)
        hs.upload(plan, th, 4321, 0)
        plan.fftp1d_real_to_complex_z(th)
        plan.s_imposebc(th)
    elif solver == "mhd":
        a, ph = [fld(10 + i) for i in range(3)], fld(13)
        for c in range(3):
            hs.upload(plan, a[c], 9876, c)
            plan.fftp1d_real_to_complex_z(a[c])
        plan.a_imposebc_and_project(a[0], a[1], a[2], ph)
    plan.synchronize()
    plan.release_scratch()      # the set-up went through the per-operator entries: give their temporaries back
    del hs


def final_state_check(plan, solver, after_steps):
    """Size-independent properties of the device state at the full size of the workload: finite positive energy,
    solenoidal field, vanishing wall-normal velocity, small slip at the walls.  Never raises."""
    import math
    try:
        fld = {"hd": plan.hd_field, "bouss": plan.bouss_field, "mhd": plan.mhd_field}[solver]
        v = [fld(i) for i in range(3)]
        div, vt0, vtL, vn0, vnL = plan.vdiagnostic(*v)
        eng = plan.energy(*v, 1)
        plan.synchronize()
        plan.release_scratch()
        out = {"after_steps": after_steps, "energy": eng, "divergence_over_energy": div / eng,
               "wall_normal_over_energy": max(vn0, vnL) / eng, "wall_tangential_over_energy": max(vt0, vtL) / eng,
               "thresholds": {"divergence": 1e-4, "wall_normal": 1e-20, "wall_tangential": 1e-2},
               "what": "vdiagnostic / energy(kin=1) on the device state (vboundary.f90:214-264, pseudospec_hd.f90:405-635)"}
        out["ok"] = bool(math.isfinite(eng) and eng > 0 and div < 1e-4 * eng and max(vn0, vnL) < 1e-20 * eng
                         and max(vt0, vtL) < 1e-2 * eng)
        return out
    except Exception as e:
        return {"ok": None, "error": f"{type(e).__name__}: {e}"}


def _oracle_case(solver, g):
    from oracle import specter_oracle as O
    if solver == "hd":
        return O.make_hd_state(g)
    if solver == "bouss":
        return O.make_bouss_state(g)
    return O.make_mhd_state(g)


def parity_check(world, rank, local, dist):
    """Pre-timing check, on EVERY run and rank count: one RK2 step of HD, BOUSS and MHD on 64^3 through the same
    fused slab-parallel path (NCCL / peer-to-peer exchange when world > 1) against the single-rank oracle, each rank
    comparing its own kx slab; the maximum over the ranks is reported.  Tolerance 1e-11 relative to the field maximum
    (north star); theta is compared on the physical rows of the mixed domain (tests/parity_cases.py)."""
    import numpy as np
    import torch
    from oracle import specter_oracle as O   # checker only
    from specter_b200 import api
    n = (64, 64, 64)
    out = {"grid": list(n), "ranks": world, "tol": 1e-11, "oracle": "oracle/specter_oracle.py (single rank)"}
    g = O.Grid(*n, CZ, OZ, Lx=1.0, Ly=0.5, Lz=1.0, tdir=TABLES, ord=2)
    nph = n[2] - CZ
    for solver in ("hd", "bouss", "mhd"):
        p = api.Plan(*n, CZ, OZ, ord=2, Lx=1.0, Ly=0.5, Lz=1.0, tdir=TABLES, nprocs=world, myrank=rank, device=local)
        if world > 1:
            p.init_comm_torch(dist, p2p_fields={"hd": (6, 3), "bouss": (8, 4), "mhd": (12, 6)}[solver])
        sl = slice(p.ista - 1, p.iend)
        cut = lambda arrs: [np.ascontiguousarray(a[sl]) for a in arrs]
        s = _oracle_case(solver, g)
        if solver == "hd":
            p.hd_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)))
            p.hd_step(1e-3, 1e-3)
            got = p.hd_get_state()[:3]
            O.hd_step(g, s, 1e-3, 1e-3)
            ref = [s.vx, s.vy, s.vz]
        elif solver == "bouss":
            p.bouss_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)))
            p.bouss_step(1e-3, 1e-3, 1e-3)
            st = p.bouss_get_state()
            O.bouss_step(g, s, 1e-3, 1e-3, 1e-3)
            got = st[:3] + [np.fft.ifft(st[4], axis=2)[:, :, :nph]]
            ref = [s.vx, s.vy, s.vz, np.fft.ifft(s.th, axis=2)[:, :, :nph]]
        else:
            p.mhd_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.fx, s.fy, s.fz, s.mx, s.my, s.mz)))
            p.mhd_step(1e-3, 1e-3, 5e-3)
            st = p.mhd_get_state()
            O.mhd_step(g, s, 1e-3, 1e-3, 5e-3)
            got = st[:3] + st[4:7]
            ref = [s.vx, s.vy, s.vz, s.ax, s.ay, s.az]
        err = 0.0
        for lo, hi in ((0, 3), (3, len(ref))):
            if hi <= lo:
                continue
            scale = max(float(np.abs(r).max()) for r in ref[lo:hi]) or 1.0
            err = max(err, max(float(np.abs(a - r[sl]).max()) for a, r in zip(got[lo:hi], ref[lo:hi])) / scale)
        if world > 1:
            t = torch.tensor([err], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            err = float(t.item())
            out.setdefault("exchanges", {})[solver] = p.comm_stats()["exchanges"]
        out[solver] = err
        p.close()
    out["ok"] = all(out[k] < out["tol"] for k in ("hd", "bouss", "mhd"))
    return out


def resolve_workload(args, world):
    """(name, nx, ny, nz, ord, dt, description, scaling) of this run: both arms use the same rule."""
    named = args.workload is not None
    wl = args.workload or ("hd512" if (args.strong or args.weak512) else DEFAULT_WORKLOAD.get(world, "hd512"))
    nx, ny, nz, ord_, dt, desc = WORKLOADS[wl]
    scaling = "weak"
    if world > 1 and not named:
        if args.weak512 or (wl == "hd512" and not args.strong):
            nx, ny = {2: (1024, 512), 4: (1024, 1024), 8: (2048, 1024)}.get(world, (512 * world, 512))
            desc = f"HD channel flow {nx}x{ny}x{nz} FP64 RK4 (512^3 points per GPU on {world} GPUs), no-slip walls, FC-Gram C=25 d=5"
        elif args.strong:
            scaling = "strong"
    elif world > 1:
        scaling = "strong"      # a named grid on N GPUs: total work fixed
    return wl, nx, ny, nz, ord_, dt, desc, scaling


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm for the path on the host cores.  The Fortran + MPI + FFTW
    binary cannot be built in this image or on the GPU box (no Fortran compiler, MPI or FFTW), so this is
    the oracle restatement (numpy + scipy.fft with an explicit worker count), kind = "port".  A "step" is ONE RK
    substep of a bounded sample grid of the configured solver; the sample actually run is named in config.workload."""
    if rank != 0:
        return
    wl, nx, ny, nz, ord_, dt, desc, _ = resolve_workload(args, max(world, args.gpus))
    solver = "bouss" if wl.startswith("bouss") else ("mhd" if wl.startswith("mhd") else "hd")
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # bounded sample: one RK substep per "step" on a grid sized so that K+W steps end within ~3 minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    cost = {"hd": 1.0, "bouss": 1.35, "mhd": 1.8}[solver]
    est = 7.5 * 8.0 / min(cores, 32) * cost  # seconds per 256^3 HD substep measured on 8 cores
    sample = (64, 64, 64)
    for cand in ((512, 512, 512), (256, 256, 512), (256, 256, 256), (128, 128, 256), (128, 128, 128), (64, 64, 64)):
        if cand[0] > nx or cand[1] > ny or cand[2] > nz:
            continue
        sample = cand
        if est * (cand[0] * cand[1] * cand[2]) / 256.0 ** 3 <= budget:
            break
    from oracle import specter_oracle as O
    O.set_workers(cores)     # scipy.fft workers set explicitly: torchrun's OMP_NUM_THREADS=1 does not apply to them
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    g = O.Grid(*sample, CZ, OZ, tdir=TABLES, ord=ord_)
    s = _oracle_case(solver, g)
    if solver == "hd":
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy()]
        sub = lambda o: O.hd_rkstep2(g, s, *C, o, dt, NU)
    elif solver == "bouss":
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy(), s.th.copy()]
        sub = lambda o: O.bouss_rkstep2(g, s, *C, o, dt, NU, KAPPA)
    else:
        C = [s.vx.copy(), s.vy.copy(), s.vz.copy(), s.ax.copy(), s.ay.copy(), s.az.copy()]
        sub = lambda o: O.mhd_rkstep2(g, s, *C, o, dt, NU, MU)
    for _ in range(args.warmup):
        sub(ord_)
    t0 = time.perf_counter()
    for k in range(args.steps):
        sub(ord_ - (k % ord_))
    t = time.perf_counter() - t0
    pts = sample[0] * sample[1] * sample[2]
    value = pts * args.steps / t
    same = list(sample) == [nx, ny, nz]
    sample_s = (f"1 RK substep of {solver.upper()} {sample[0]}x{sample[1]}x{sample[2]} per step (oracle restatement, numpy + scipy.fft "
                f"workers={cores}); throughput is per grid point, the configured grid is {nx}x{ny}x{nz}")
    line = {"impl": "reference", "metric": "grid-point RK substeps per second", "value": value, "unit": "pts*substep/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (desc if same else f"SAMPLE {sample[0]}x{sample[1]}x{sample[2]} of: {desc}"),
                       "grid": [nx, ny, nz], "sample_grid": list(sample), "sample_is_full_grid": same,
                       "rk_order": ord_, "Cz": CZ, "oz": OZ, "solver": solver},
            "cpu_baseline": {"value": value, "unit": "pts*substep/s", "cores": cores, "kind": "port", "sample": sample_s},
            "e2e": {"value": value, "unit": "pts*substep/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))   # the process's real stdout (fd 1 was pointed at stderr by quiet_stdout)


def cpu_oracle_substep_rate(nx, ny, nz, ord_, dt, reps=1, workers=None):
    """The oracle (numpy/scipy restatement of the reference's pass structure) timed on the host cores."""
    from oracle import specter_oracle as O
    if workers:
        O.set_workers(workers)
    g = O.Grid(nx, ny, nz, CZ, OZ, tdir=TABLES, ord=ord_)
    s = O.make_hd_state(g)
    C = [s.vx.copy(), s.vy.copy(), s.vz.copy()]
    O.hd_rkstep2(g, s, *C, ord_, dt, NU)  # warm-up (FFT plans, page faults)
    t0 = time.perf_counter()
    for r in range(reps):
        O.hd_rkstep2(g, s, *C, max(ord_ - 1 - r, 1), dt, NU)
    t = (time.perf_counter() - t0) / reps
    return nx * ny * nz / t, t


_REAL_STDOUT = None


def emit(text):
    """The one JSON line goes to the process's original stdout; everything else a library prints on fd 1
    (e.g. NCCL's version banner) was redirected to stderr by quiet_stdout()."""
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: hd512 on 1 GPU, the per-GPU problem of hd2048 on 2 / 4 GPUs, hd2048 on 8")
    ap.add_argument("--path", type=int, default=0, help="0 = fused substep (product default), 1 = per-operator composition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the 64^3 HD / BOUSS / MHD parity check before the timing")
    ap.add_argument("--no-state-check", action="store_true", help="skip the full-size property check of the final state (1 GPU)")
    ap.add_argument("--strong", action="store_true", help="N > 1 without --workload: keep the 512^3 grid (strong scaling)")
    ap.add_argument("--weak512", action="store_true", help="N > 1 without --workload: 512^3 points per GPU (round-1 weak-scaling line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    quiet_stdout()
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from specter_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: specter_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl, nx, ny, nz, ord_, dt, desc, scaling = resolve_workload(args, world)
    solver = "bouss" if wl.startswith("bouss") else ("mhd" if wl.startswith("mhd") else "hd")
    b_alg = B_ALG_BY_SOLVER[solver]

    parity = None
    if not args.no_parity:
        parity = parity_check(world, rank, local, dist)

    plan = api.Plan(nx, ny, nz, CZ, OZ, ord=ord_, tdir=TABLES, nprocs=world, myrank=rank, device=local)
    if world > 1:
        plan.init_comm_torch(dist, p2p_fields={"hd": (6, 3), "bouss": (8, 4), "mhd": (12, 6)}[solver])
    device_state(plan, solver)
    if solver == "hd":
        step = lambda: plan.hd_step(dt, NU, impl=args.path)
    elif solver == "bouss":
        step = lambda: plan.bouss_step(dt, NU, KAPPA, impl=args.path)
    else:
        step = lambda: plan.mhd_step(dt, NU, MU, impl=args.path)
    if solver != "hd":
        args.no_e2e = True          # the host-buffer entry exists for the headline (HD) path
        args.no_cpu_baseline = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plan.synchronize()

    # end-to-end inputs: the synthetic state as pinned HOST arrays (v, p': 4 fields per rank), read back from the device
    # before the timed steps change it.  Skipped (and said so) when the host cannot page-lock that much.
    pinned, e2e_skip = None, None
    if not args.no_e2e:
        need = 4 * 16 * int(np.prod(plan.cshape)) * world
        try:
            with open("/proc/meminfo") as fh:
                avail = next(int(ln.split()[1]) * 1024 for ln in fh if ln.startswith("MemAvailable"))
        except Exception:
            avail = None
        if avail is not None and need > 0.6 * avail:
            e2e_skip = f"host has {avail / 2**30:.0f} GiB available, the pinned state of {world} ranks needs {need / 2**30:.0f} GiB"
        else:
            try:
                pinned = [plan.pinned_empty() for _ in range(4)]
                plan._call("sx_hd_get_state", *[a.ctypes.data for a in pinned])
            except Exception as e:      # e.g. not enough page-locked host memory for the largest grids
                pinned, e2e_skip = None, f"{type(e).__name__}: {e}"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    n0 = plan.launch_count
    barrier()
    sampler.mark_begin()
    plan.time_begin()
    for _ in range(args.steps):
        step()
    ms = plan.time_end()
    sampler.mark_end()
    barrier()
    launches = plan.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    npts = float(nx) * ny * nz
    value = npts * ord_ * args.steps / (ms * 1e-3)
    peak, peak_src = peaks()

    # ---- per-kernel device times (a separate, untimed-for-the-headline pass with stage events) ----
    if world > 1:
        plan.comm_stats(reset=True)
    plan.stage_timing(True)
    for _ in range(2):
        step()
    stages = plan.stage_times()
    plan.stage_timing(False)
    comm = plan.comm_stats() if world > 1 else None
    stage_report, dominant = {}, None
    tot_ms = sum(v[0] for v in stages.values()) or 1.0
    model = stage_bytes_per_pt((nz - CZ) / nz, solver)
    for name, (sms, cnt) in stages.items():
        per_launch = sms / cnt
        launches_per_substep = cnt / (2.0 * ord_)
        bpp = model.get(name)
        entry = {"ms_per_launch": per_launch, "launches_per_substep": launches_per_substep, "share": sms / tot_ms}
        if bpp:
            alg_bytes = bpp * npts / world / launches_per_substep
            entry["algorithmic_bytes_per_launch"] = alg_bytes
            entry["achieved_gbs"] = alg_bytes / (per_launch * 1e-3) / 1e9
            entry["frac_of_hbm_peak"] = entry["achieved_gbs"] / peak
            if dominant is None or sms > stages[dominant][0]:
                dominant = name
        stage_report[name] = entry
    dom = stage_report.get(dominant, {})
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this same command
    # (dram__bytes_read.sum + dram__bytes_write.sum; profiles/ncu_traffic.json, written by tools/ncu_summary.py); the
    # capture carries the hash of the kernel sources it was taken from and is dropped when that is not this code
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        ent = tj.get("workloads", {}).get(wl) if "workloads" in tj else (tj if tj.get("workload") == wl else None)
        if ent and world == 1:
            cur = sources_hash()
            traffic_src = {"capture": ent.get("capture"), "sources_hash": ent.get("sources_hash"), "current_sources_hash": cur,
                           "stale": ent.get("sources_hash") != cur}
            if not traffic_src["stale"]:
                traffic = ent["stages"].get(dominant, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError, KeyError):
        traffic = None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": dom.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                "frac": dom.get("frac_of_hbm_peak"), "traffic": traffic, "traffic_from_profile": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom.get("algorithmic_bytes_per_launch"),
                "avg_launch_ms": dom.get("ms_per_launch"),
                "whole_substep": {"algorithmic_bytes_per_point": b_alg,
                                  "achieved": b_alg * value / world / 1e9, "frac": b_alg * value / world / 1e9 / peak,
                                  "frac_of_nominal_8000_gbs": b_alg * value / world / 1e9 / 8000.0}}   # both peaks

    # ---- size-independent properties of the state the timed steps produced, at the FULL size of the workload (the oracle
    # never runs there): finite energy, solenoidal field, vanishing wall-normal velocity, small slip at the walls.  One
    # GPU only (the 8-GPU workload leaves no room for the diagnostics' temporaries); never fatal for the bench line.
    state_check = None
    if world == 1 and not args.no_state_check:
        state_check = final_state_check(plan, solver, args.warmup + args.steps + 2)

    # ---- end to end through the host-buffer C-ABI entry (H2D + ord substeps + D2H per step) ----
    e2e = None
    if not args.no_e2e and pinned is None:
        e2e = {"value": None, "skipped": e2e_skip}
    elif not args.no_e2e:
        ksteps = max(2, min(args.steps, 3))
        fb = pinned[0].nbytes
        # the (constant) forcing is resident on the device since the set-up; the calls pass NULL for it
        plan.hd_step_host(*pinned, None, None, None, dt, NU)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            plan.hd_step_host(*pinned, None, None, None, dt, NU)
        barrier()
        te = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        wall_rows = 2 * 16 * ny * (nx // 2 + 1)   # the two wall rows of p' (all the step reads of it), summed over the ranks
        e2e = {"value": npts * ord_ * ksteps / te, "unit": "pts*substep/s", "h2d_bytes_per_step": 3 * fb * world + wall_rows,
               "d2h_bytes_per_step": 4 * fb * world, "steps": ksteps, "ms_per_step": 1e3 * te / ksteps,
               "api": "sx_hd_step_host (pinned host arrays in the reference layout; v up, the wall rows of p' up -- the step reads "
                      "nothing else of it --, v and p' down every step; the constant body force uploaded once and kept resident)"}
        e2e["finite"] = all(bool(np.isfinite(a).all()) for a in pinned[:4])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sn = (256, 256, 256) if nx >= 256 else (nx, ny, nz)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        rate, secs = cpu_oracle_substep_rate(*sn, ord_, dt, reps=2, workers=cores)
        cpu = {"value": rate, "unit": "pts*substep/s", "cores": cores, "kind": "port",
               "sample": f"2 RK substeps of HD {sn[0]}x{sn[1]}x{sn[2]} ({secs:.1f} s each) with the numpy/scipy.fft oracle "
                         "restatement (not the reference MPI+OpenMP+FFTW binary, which cannot be built here)"}

    if rank == 0:
        line = {"metric": "grid-point RK substeps per second", "value": value, "unit": "pts*substep/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "ms_per_substep": ms / args.steps / ord_, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "name": wl, "solver": solver, "grid": [nx, ny, nz], "points_per_gpu": npts / world,
                           "rk_order": ord_, "Cz": CZ, "oz": OZ, "dt": dt, "nu": NU,
                           "path": "fused" if args.path == 0 else "per-operator",
                           "l2": "inputs larger than L2 (each pass streams >= 3 GB; L2 = 126 MB)",
                           "parallelism": f"slab x{world}", "exchange": ("peer-to-peer copies + NCCL barrier" if getattr(plan, "p2p", False) else "NCCL send/recv") if world > 1 else "none"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "parity_check": parity, "state_check": state_check, "stages": stage_report}
        if comm and comm["exchanges"] and comm["ms"] > 0:
            # NVLink roofline: bytes this rank sent / device time of the exchange rounds on the communication stream in the
            # stage-timing pass (event pairs resolved after the pass: the exchange overlaps the compute stream as in the
            # timed run, so this is the rate under HBM contention, not an isolated copy)
            per_ex = comm["bytes_sent"] / comm["exchanges"]
            ms_ex = comm["ms"] / comm["exchanges"]
            line["nvlink"] = {"bytes_sent_per_exchange": per_ex, "ms_per_exchange": ms_ex,
                              "exchanges_per_substep": comm["exchanges"] / (2.0 * ord_),
                              "bytes_sent_per_substep": comm["bytes_sent"] / (2.0 * ord_),
                              "exchange_stream_ms_per_substep": comm["ms"] / (2.0 * ord_),
                              "achieved_gbs_per_direction": per_ex / (ms_ex * 1e-3) / 1e9,
                              "peak_gbs_per_direction": 900.0, "frac": per_ex / (ms_ex * 1e-3) / 1e9 / 900.0}
        emit(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
