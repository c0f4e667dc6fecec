"""bench.py's contract where it can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads; only rank 0 of a multi-rank launch does the work; our arm refuses to run without a CUDA device (no CPU
fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "hd64"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pts*substep/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 64 ** 3 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_exit_without_work():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "hd64"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "3", "--workload", "hd64", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workload_rule_and_byte_models():
    """The workload both arms run for a rank count (BASELINE configs[1] on 1 GPU, 512^3 per GPU on 2 / 4, configs[4] on 8)
    and the per-kernel byte models: their HD sum is what the kernels move as built, never less than the
    440 B per point of SURVEY 8(d) that the whole-substep fraction is quoted on."""
    import argparse
    import sys
    sys.path.insert(0, ROOT)
    import bench
    ns = lambda **kw: argparse.Namespace(**{"workload": None, "strong": False, "weak512": False, **kw})
    assert bench.resolve_workload(ns(), 1)[:4] == ("hd512", 512, 512, 512)
    assert bench.resolve_workload(ns(), 2)[1:4] == (1024, 512, 512) and bench.resolve_workload(ns(), 2)[7] == "weak"
    assert bench.resolve_workload(ns(), 4)[1:4] == (1024, 1024, 512)
    assert bench.resolve_workload(ns(), 8)[:4] == ("hd2048", 2048, 2048, 1024)
    assert bench.resolve_workload(ns(strong=True), 4)[1:4] == (512, 512, 512) and bench.resolve_workload(ns(strong=True), 4)[7] == "strong"
    assert bench.resolve_workload(ns(workload="mhd512"), 8)[7] == "strong"
    r = (512 - 25) / 512
    hd = bench.stage_bytes_per_pt(r, "hd")
    assert abs(sum(hd.values()) / 8 - (22 + 42 * r)) < 1e-9          # 62 F as built
    assert sum(hd.values()) >= bench.B_ALG_BY_SOLVER["hd"]
    for solver in ("bouss", "mhd"):
        assert sum(bench.stage_bytes_per_pt(r, solver).values()) >= bench.B_ALG_BY_SOLVER[solver] * 0.95
    assert len(bench.sources_hash()) == 12


def test_final_state_check_on_the_emulation(emu_lib, tables):
    """bench.py's full-size property check of the final state (state_check in the JSON line), here on a small grid through
    the emulation build: the synthetic state after a step is solenoidal, has no wall-normal velocity and finite energy;
    a failing call is reported, never raised."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from specter_b200 import api
    for solver in ("hd", "bouss", "mhd"):
        p = api.Plan(32, 16, 64, 25, 5, ord=2, tdir=tables, lib=emu_lib)
        bench.device_state(p, solver)
        {"hd": lambda: p.hd_step(1e-3, 1e-3), "bouss": lambda: p.bouss_step(1e-3, 1e-3, 1e-3),
         "mhd": lambda: p.mhd_step(1e-3, 1e-3, 5e-3)}[solver]()
        sc = bench.final_state_check(p, solver, 1)
        assert sc["ok"] is True, sc
        p.close()
    assert bench.final_state_check(None, "hd", 0)["ok"] is None


def test_traffic_capture_belongs_to_these_kernel_sources():
    """roofline.traffic comes from the committed ncu capture (profiles/ncu_traffic.json) and is dropped when the kernel
    sources differ from the captured ones: the committed state must carry a capture of ITS sources.  The hash is taken over
    the GPU view of the files -- what only the test emulation compiles does not count."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
        tj = json.load(fh)
    cur = bench.sources_hash()
    assert {w: e["sources_hash"] for w, e in tj["workloads"].items()} == {w: cur for w in tj["workloads"]}
    src = "a\n#ifndef SX_EMU\ngpu\n#if X\nnested\n#endif\n#else\nemu\n#endif\n#ifdef SX_EMU\nemu2\n#endif\nb\n"
    assert bench.gpu_view(src) == "a\ngpu\n#if X\nnested\n#endif\nb\n"


def test_bench_main_end_to_end_on_the_emulation(emu_lib):
    """bench.py's own arm cannot run here (no GPU, and it refuses to: test_our_arm_needs_a_gpu).  This runs its WHOLE main()
    in a child process in which -- from the test side, bench.py has no such switch -- torch.cuda's presence check is patched
    and the product loader hands out the CPU-thread emulation of the kernel sources: set-up of the synthetic state, warm-up
    and timed steps, the stage-timing pass, the per-kernel roofline arithmetic, the full-size state check, the host-buffer
    (e2e) leg on page-locked arrays, the CPU-baseline leg and the one JSON line with every key the driver reads.  The
    numbers are meaningless; the point is that no line of the path the driver runs at round end is executed for the
    first time on the GPU box."""
    code = ("import sys, runpy\n"
            "sys.path.insert(0, %r)\n"
            "import torch\n"
            "torch.cuda.is_available = lambda: True\n"
            "torch.cuda.set_device = lambda *a, **k: None\n"
            "torch.cuda.synchronize = lambda *a, **k: None\n"
            "from specter_b200 import api\n"
            "emu = api.Library(%r)\n"
            "api.load_library = lambda *a, **k: emu\n"
            "sys.argv = ['bench.py', '--workload', 'hd64', '--steps', '2', '--warmup', '1', '--no-parity']\n"
            "runpy.run_path(%r, run_name='__main__')\n") % (ROOT, emu_lib.path, BENCH)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "state_check", "stages"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["value"] > 0   # W >= 3 is enforced
    assert d["config"]["name"] == "hd64" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["kernel"] in d["stages"] and rf["frac"] > 0 and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["whole_substep"]["algorithmic_bytes_per_point"] == 440.0
    e = d["e2e"]
    assert e["value"] > 0 and e["finite"] is True and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    assert d["state_check"]["ok"] is True, d["state_check"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert abs(d["value"] - 64 ** 3 * 2 * 2 / (d["ms_per_step"] * 2 * 1e-3)) < 1e-6 * d["value"]
