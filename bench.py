#!/usr/bin/env python
"""Benchmark of the per-RK-substep hot path (BASELINE.json metric: grid-point RK substeps per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload hd512|hd64|...]

A "step" is one full Runge-Kutta time step (rkstep1 + `ord` substeps, specter.fpp:1142-1161) of the HD
solver on synthetic initial conditions; value = nx*ny*nz*ord*K / device seconds (continuation planes
included, as the reference's benchmark.txt counts them).  Default workload = BASELINE.json configs[1]:
HD channel flow 512^3, FP64, RK4, one B200.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
    "hd512": (512, 512, 512, 4, 2e-4, "HD channel flow 512x512x512 FP64 RK4, no-slip walls, FC-Gram C=25 d=5"),
    "hd256": (256, 256, 256, 4, 5e-4, "HD channel flow 256^3 FP64 RK4 (reduced; not the headline config)"),
    "hd64": (64, 64, 64, 2, 1e-3, "HD 64^3 RK2 (BASELINE configs[0], parity config)"),
    "hd1024": (1024, 1024, 512, 4, 1e-4, "HD 1024x1024x512 FP64 RK4 (multi-GPU sizes; needs >= 4 GPUs)"),
    "hd2048": (2048, 2048, 1024, 4, 5e-5, "HD 2048x2048x1024 FP64 RK4 (BASELINE configs[4]; needs 8 GPUs)"),
    "bouss512": (512, 512, 512, 4, 2e-4, "BOUSS Rayleigh-Benard 512x512x512 FP64 RK4, no-slip + constant-temperature walls"),
    "bouss1024": (1024, 1024, 512, 4, 1e-4, "BOUSS Rayleigh-Benard 1024x1024x512 FP64 RK4 (BASELINE configs[2]; needs 8 GPUs)"),
    "mhd512": (512, 512, 512, 4, 2e-4, "MHD vector potential 512x512x512 FP64 RK4, no-slip + conducting walls (BASELINE configs[3])"),
}
CZ, OZ, NU = 25, 5, 1e-3
KAPPA, MU = 1e-3, 5e-3
# algorithmic HBM bytes per grid-point-substep (55 F / 73 F / 98 F, F = 8 B/pt)
B_ALG_BY_SOLVER = {"hd": 440.0, "bouss": 584.0, "mhd": 784.0}
B_ALG = 440.0
# algorithmic bytes per grid point of each pass (DESIGN.md "Kernels"): F = 8 B per point per full-field read or write
def stage_bytes_per_pt(r):
    """r = physical rows / nz: only physical rows cross the transposition."""
    F = 8.0
    return {"zinv_tile": (3 + 6 * r) * F, "yinv_tile": 15 * r * F, "xpass": 12 * r * F, "yfwd_tile": 6 * r * F,
            "zfwd_rk": (12 + 3 * r) * F, "project": 7 * F}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/sx_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_state(plan, seed=1234):
    """Synthetic initial condition built with the product API only: random low-wavenumber modes with a
    wall-vanishing z envelope in the mixed (z,ky,kx) domain -> continued z-FFT -> projected onto
    solenoidal no-slip fields by sx_v_imposebc_and_project; uniform body force f0=1 in x
    (initialfv.f90:25-31).  Values do not affect timing (no data-dependent control flow)."""
    import numpy as np
    nxl, ny, nz = plan.cshape
    nph = nz - plan.Cz
    rng = np.random.default_rng(seed + plan.ista)
    z = np.arange(nph) / (nph - 1.0)
    env = np.sin(np.pi * z) ** 2
    N = float(plan.nx) * plan.ny * plan.nz
    fields = []
    kmax = 4
    for c in range(3):
        a = np.zeros((nxl, ny, nz), dtype=np.complex128)
        for i in range(nxl):
            kx = plan.ista - 1 + i
            if kx > kmax:
                break
            for j in list(range(0, kmax + 1)) + list(range(ny - kmax, ny)):
                if kx == 0 and j > ny // 2:
                    continue
                amp = (rng.standard_normal() + 1j * rng.standard_normal()) * (N / plan.nz) * 0.05
                prof = env * np.cos(np.pi * (1 + (i + j + c) % 3) * z)
                a[i, j, :nph] = amp * prof
                if kx == 0:
                    if j == 0:
                        a[i, j, :nph] = (amp.real * prof)
                    else:
                        a[i, ny - j, :nph] = np.conj(a[i, j, :nph])
        fields.append(a)
    dev = [plan.spectral(a) for a in fields]
    for d in dev:
        plan.fftp1d_real_to_complex_z(d)
    pr = plan.spectral(np.zeros((nxl, ny, nz), dtype=np.complex128))
    plan.v_imposebc_and_project(dev[0], dev[1], dev[2], pr, plan.ord)
    host = [d.get() for d in dev]
    for d in dev + [pr]:
        d.free()
    f = [np.zeros((nxl, ny, nz), dtype=np.complex128) for _ in range(3)]
    if plan.ista == 1:
        f[0][0, 0, 0] = 1.0 * N
    return host + [np.zeros((nxl, ny, nz), dtype=np.complex128)] + f


def _low_modes(plan, seed, ncomp):
    import numpy as np
    nxl, ny, nz = plan.cshape
    nph = nz - plan.Cz
    rng = np.random.default_rng(seed + plan.ista)
    z = np.arange(nph) / (nph - 1.0)
    env = np.sin(np.pi * z) ** 2
    N = float(plan.nx) * plan.ny * plan.nz
    kmax = 4
    out = []
    for c in range(ncomp):
        a = np.zeros((nxl, ny, nz), dtype=np.complex128)
        for i in range(nxl):
            kx = plan.ista - 1 + i
            if kx > kmax:
                break
            for j in list(range(0, kmax + 1)) + list(range(ny - kmax, ny)):
                if kx == 0 and j > ny // 2:
                    continue
                amp = (rng.standard_normal() + 1j * rng.standard_normal()) * (N / plan.nz) * 0.05
                prof = env * np.cos(np.pi * (1 + (i + j + c) % 3) * z)
                a[i, j, :nph] = amp * prof
                if kx == 0:
                    if j == 0:
                        a[i, j, :nph] = (amp.real * prof)
                    else:
                        a[i, ny - j, :nph] = np.conj(a[i, j, :nph])
        out.append(a)
    return out


def synthetic_scalar(plan, seed=4321):
    """Temperature fluctuation with constant (zero) walls: low modes -> continued z-FFT -> sx_s_imposebc."""
    d = plan.spectral(_low_modes(plan, seed, 1)[0])
    plan.fftp1d_real_to_complex_z(d)
    plan.s_imposebc(d)
    h = d.get()
    d.free()
    return h


def synthetic_potential(plan, seed=9876):
    """Vector potential satisfying the conducting-wall conditions: low modes -> continued z-FFT ->
    sx_a_imposebc_and_project."""
    dev = [plan.spectral(a) for a in _low_modes(plan, seed, 3)]
    for d in dev:
        plan.fftp1d_real_to_complex_z(d)
    ph = plan.spectral()
    plan.a_imposebc_and_project(dev[0], dev[1], dev[2], ph)
    host = [d.get() for d in dev]
    for d in dev + [ph]:
        d.free()
    return host


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


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm for the path (the oracle restatement -- the
    Fortran+MPI+FFTW binary cannot be built in this image) on the host cores."""
    if rank != 0:
        return
    nx, ny, nz, ord_, dt, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # bounded sample: one RK substep per "step" on a grid sized so that K+W steps end within ~3 minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = (nx, ny, nz)
    est = 7.5 * 8.0 / min(cores, 32)  # seconds per 256^3 substep measured on 8 cores
    for cand in ((512, 512, 512), (256, 256, 512), (256, 256, 256), (128, 128, 256), (128, 128, 128), (64, 64, 64)):
        if cand[0] > nx or cand[2] > nz:
            continue
        sample = cand
        if est * (cand[0] * cand[1] * cand[2]) / 256.0 ** 3 <= budget:
            break
    from oracle import specter_oracle as O
    O.set_workers(cores)
    g = O.Grid(*sample, CZ, OZ, tdir=TABLES, ord=ord_)
    s = O.make_hd_state(g)
    C = [s.vx.copy(), s.vy.copy(), s.vz.copy()]
    for _ in range(args.warmup):
        O.hd_rkstep2(g, s, *C, ord_, dt, NU)
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.hd_rkstep2(g, s, *C, ord_ - (k % ord_), dt, NU)
    t = time.perf_counter() - t0
    pts = sample[0] * sample[1] * sample[2]
    value = pts * args.steps / t
    sample_s = f"1 RK substep of HD {sample[0]}x{sample[1]}x{sample[2]} per step (oracle restatement, numpy+scipy.fft workers={cores})"
    line = {"impl": "reference", "metric": "grid-point RK substeps per second", "value": value, "unit": "pts*substep/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "grid": [nx, ny, nz], "rk_order": ord_, "Cz": CZ, "oz": OZ},
            "cpu_baseline": {"value": value, "unit": "pts*substep/s", "cores": cores, "kind": "port", "sample": sample_s},
            "e2e": {"value": value, "unit": "pts*substep/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))   # the process's real stdout (fd 1 was pointed at stderr by quiet_stdout)


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
    ap.add_argument("--workload", default="hd512", choices=sorted(WORKLOADS))
    ap.add_argument("--path", type=int, default=0, help="0 = fused substep (product default), 1 = per-operator composition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strong", action="store_true", help="N > 1: keep the 512^3 grid (strong scaling) instead of 512^3 points per GPU")
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
    nx, ny, nz, ord_, dt, desc = WORKLOADS[args.workload]
    weak = world > 1 and args.workload == "hd512" and not args.strong
    if weak:
        # the headline workload per GPU: 512^3 points each, the periodic directions grow with the rank count
        nx, ny = {2: (1024, 512), 4: (1024, 1024), 8: (2048, 1024)}.get(world, (512 * world, 512))
        desc = f"HD channel flow {nx}x{ny}x{nz} FP64 RK4 (512^3 points per GPU on {world} GPUs), no-slip walls, FC-Gram C=25 d=5"
    solver = "bouss" if args.workload.startswith("bouss") else ("mhd" if args.workload.startswith("mhd") else "hd")
    b_alg = B_ALG_BY_SOLVER[solver]
    plan = api.Plan(nx, ny, nz, CZ, OZ, ord=ord_, tdir=TABLES, nprocs=world, myrank=rank, device=local)
    if world > 1:
        plan.init_comm_torch(dist, p2p_fields={"hd": (6, 3), "bouss": (8, 4), "mhd": (12, 6)}[solver])
    st = synthetic_state(plan)
    if args.workload in ("hd1024", "hd2048", "bouss1024"):
        plan.release_scratch()      # the set-up went through the per-operator entries: give their temporaries back (6 fields)
    zero = np.zeros_like(st[0])
    if solver == "hd":
        plan.hd_put_state(*st)
        step = lambda: plan.hd_step(dt, NU, impl=args.path)
    elif solver == "bouss":
        th = synthetic_scalar(plan)
        plan.bouss_put_state(st[0], st[1], st[2], st[3], th, st[4], st[5], st[6], zero)
        step = lambda: plan.bouss_step(dt, NU, KAPPA, impl=args.path)
    else:
        a = synthetic_potential(plan)
        plan.mhd_put_state(st[0], st[1], st[2], st[3], a[0], a[1], a[2], zero, zero, zero, zero, zero, zero)
        step = lambda: plan.mhd_step(dt, NU, MU, impl=args.path)
    if solver != "hd":
        args.no_e2e = True          # the host-buffer entry exists for the headline (HD) path
        args.no_cpu_baseline = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plan.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = plan.launch_count
    barrier()
    plan.time_begin()
    for _ in range(args.steps):
        step()
    ms = plan.time_end()
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
    for name, (sms, cnt) in stages.items():
        per_launch = sms / cnt
        launches_per_substep = cnt / (2.0 * ord_)
        bpp = stage_bytes_per_pt((nz - CZ) / nz).get(name) if solver == "hd" else None
        entry = {"ms_per_launch": per_launch, "launches_per_substep": launches_per_substep, "share": sms / tot_ms}
        if bpp:
            alg_bytes = bpp * npts / world / launches_per_substep
            entry["algorithmic_bytes_per_launch"] = alg_bytes
            entry["achieved_gbs"] = alg_bytes / (per_launch * 1e-3) / 1e9
            entry["frac_of_hbm_peak"] = entry["achieved_gbs"] / peak
        stage_report[name] = entry
        if dominant is None or sms > stages[dominant][0]:
            dominant = name
    dom = stage_report.get(dominant, {})
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this same command
    # (dram__bytes_read.sum + dram__bytes_write.sum; profiles/ncu_traffic.json, written by tools/ncu_summary.py)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj["stages"].get(dominant, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError, KeyError):
        traffic = None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": dom.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                "frac": dom.get("frac_of_hbm_peak"), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom.get("algorithmic_bytes_per_launch"),
                "avg_launch_ms": dom.get("ms_per_launch"),
                "whole_substep": {"algorithmic_bytes_per_point": b_alg,
                                  "achieved": b_alg * value / world / 1e9, "frac": b_alg * value / world / 1e9 / peak,
                                  "frac_of_nominal_8000_gbs": b_alg * value / world / 1e9 / 8000.0}}   # both peaks
    if solver != "hd":   # per-kernel byte model exists for the HD kernels only: report the whole substep
        roofline.update(kernel="whole substep", achieved=roofline["whole_substep"]["achieved"],
                        frac=roofline["whole_substep"]["frac"], algorithmic_bytes_per_launch=b_alg * npts / world,
                        avg_launch_ms=ms / args.steps / ord_)

    # ---- end to end through the host-buffer C-ABI entry (H2D + ord substeps + D2H per step) ----
    e2e = None
    if not args.no_e2e:
        pinned = [plan.pinned_like(a) for a in st]
        ksteps = max(2, min(args.steps, 3))
        plan.hd_step_host(*pinned, dt, NU)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            plan.hd_step_host(*pinned, dt, NU)
        barrier()
        te = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        fb = st[0].nbytes
        e2e = {"value": npts * ord_ * ksteps / te, "unit": "pts*substep/s", "h2d_bytes_per_step": 7 * fb * world,
               "d2h_bytes_per_step": 4 * fb * world, "steps": ksteps, "ms_per_step": 1e3 * te / ksteps,
               "api": "sx_hd_step_host (pinned host arrays in the reference layout)"}
        ok = all(bool(np.isfinite(a).all()) for a in pinned[:4])
        e2e["finite"] = ok

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sn = (256, 256, 256) if nx >= 256 else (nx, ny, nz)
        rate, secs = cpu_oracle_substep_rate(*sn, ord_, dt, reps=2)
        cpu = {"value": rate, "unit": "pts*substep/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"2 RK substeps of HD {sn[0]}x{sn[1]}x{sn[2]} ({secs:.1f} s each) with the numpy/scipy.fft oracle "
                         "restatement (not the reference MPI+OpenMP+FFTW binary, which cannot be built here)"}

    if rank == 0:
        line = {"metric": "grid-point RK substeps per second", "value": value, "unit": "pts*substep/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "ms_per_substep": ms / args.steps / ord_, "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "grid": [nx, ny, nz], "rk_order": ord_, "Cz": CZ, "oz": OZ, "dt": dt, "nu": NU,
                           "path": "fused" if args.path == 0 else "per-operator",
                           "l2": "inputs larger than L2 (each pass streams >= 3 GB; L2 = 126 MB)",
                           "parallelism": f"slab x{world}", "exchange": ("peer-to-peer copies + NCCL barrier" if getattr(plan, "p2p", False) else "NCCL send/recv") if world > 1 else "none"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "stages": stage_report}
        if comm and comm["exchanges"] and comm["ms"] > 0:
            # NVLink roofline: bytes this rank sent per exchange / device time of the exchange on the comm
            # stream, measured un-overlapped in the stage-timing pass (sx_plan_comm_stats)
            per_ex = comm["bytes_sent"] / comm["exchanges"]
            ms_ex = comm["ms"] / comm["exchanges"]
            line["nvlink"] = {"bytes_sent_per_exchange": per_ex, "ms_per_exchange": ms_ex,
                              "exchanges_per_substep": comm["exchanges"] / (2.0 * ord_),
                              "achieved_gbs_per_direction": per_ex / (ms_ex * 1e-3) / 1e9,
                              "peak_gbs_per_direction": 900.0, "frac": per_ex / (ms_ex * 1e-3) / 1e9 / 900.0}
        emit(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
