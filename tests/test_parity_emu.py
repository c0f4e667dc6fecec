"""CPU (no GPU): the kernel SOURCES under the CPU-thread emulation vs the oracle, small grids.
Checks index math, barrier structure and the host-side composition; the GPU parity proper is
test_parity_gpu.py."""
import pytest

import parity_cases as P

SMALL = (32, 16, 64)


def test_fft1d_z(emu_lib, tables):
    P.case_fft1d_z(emu_lib, tables, SMALL)
    P.case_fft1d_z(emu_lib, tables, (16, 16, 128), Cz=25)
    P.case_fft1d_z(emu_lib, tables, (16, 16, 32), Cz=0)


def test_normalisations(emu_lib, tables):
    P.case_normalisations(emu_lib, tables, SMALL)


def test_goto_domain(emu_lib, tables):
    P.case_goto_domain(emu_lib, tables, SMALL)


def test_fft3d(emu_lib, tables):
    P.case_fft3d(emu_lib, tables, SMALL)
    P.case_fft3d(emu_lib, tables, (16, 32, 64))
    P.case_fft3d(emu_lib, tables, (64, 16, 16), Cz=0)


def test_fft_known_answer(emu_lib, tables):
    P.case_fft_known_answer(emu_lib, tables, 32)


def test_spectral_ops(emu_lib, tables):
    P.case_spectral_ops(emu_lib, tables, SMALL)


def test_nonlinear(emu_lib, tables):
    P.case_nonlinear(emu_lib, tables, SMALL)


def test_projection(emu_lib, tables):
    P.case_projection(emu_lib, tables, SMALL)


def test_diagnostics(emu_lib, tables):
    P.case_diagnostics(emu_lib, tables, SMALL)


@pytest.mark.parametrize("impl", [1, 0])
def test_hd_substeps(emu_lib, tables, impl):
    P.case_hd_substeps(emu_lib, tables, SMALL, ord=2, nsteps=2, impl=impl)


def test_hd_substeps_rk4_moving_walls(emu_lib, tables):
    P.case_hd_substeps(emu_lib, tables, (16, 16, 64), ord=4, nsteps=1, impl=0, walls=((0.2, -0.1), (-0.3, 0.1)))


def test_hd_step_host(emu_lib, tables):
    P.case_hd_step_host(emu_lib, tables, SMALL)
    P.case_hd_step_host(emu_lib, tables, (16, 16, 64), pinned=True, nsteps=2, inflight=True)


def test_advect_vector(emu_lib, tables):
    P.case_advect_vector(emu_lib, tables, SMALL)


def test_scalar_vecpot_bc(emu_lib, tables):
    P.case_scalar_vecpot_bc(emu_lib, tables, SMALL)


@pytest.mark.parametrize("impl", [1, 0])
def test_bouss_substeps(emu_lib, tables, impl):
    P.case_bouss_substeps(emu_lib, tables, SMALL, ord=2, nsteps=1, impl=impl)


@pytest.mark.parametrize("impl", [1, 0])
def test_mhd_substeps(emu_lib, tables, impl):
    P.case_mhd_substeps(emu_lib, tables, SMALL, ord=2, nsteps=1, impl=impl)


def test_hd_substeps_bulk_xpass(emu_lib, tables, monkeypatch):
    # the bulk-copy (TMA ring) x pass is the default from nx = 256; force it on a small grid
    monkeypatch.setenv("SX_XP", "10")
    P.case_hd_substeps(emu_lib, tables, (64, 16, 64), ord=2, nsteps=1, impl=1)


def test_bouss_substeps_bulk_xpass(emu_lib, tables, monkeypatch):
    monkeypatch.setenv("SX_XP", "10")
    P.case_bouss_substeps(emu_lib, tables, (64, 16, 64), ord=2, nsteps=1, impl=1)


def test_hd_substeps_bulk_tiles(emu_lib, tables, monkeypatch):
    # the bulk-copy (TMA) tile kernels are the default from length 256; force them on a small grid
    monkeypatch.setenv("SX_TMA_MIN", "16")
    P.case_hd_substeps(emu_lib, tables, (16, 128, 128), ord=2, nsteps=1, impl=0)


def test_hd_substeps_bulk_project(emu_lib, tables, monkeypatch):
    # bulk-copy projection kernel (default from nz = 256): wall rows by reduction, exponentials by recurrence
    monkeypatch.setenv("SX_PJ", "10")
    P.case_hd_substeps(emu_lib, tables, (16, 16, 256), ord=2, nsteps=2, impl=0, walls=((0.2, -0.1), (-0.3, 0.1)))


def test_mhd_substeps_bulk_kernels(emu_lib, tables, monkeypatch):
    # the cross-product x pass on the bulk-copy ring (forced on a short line) and, from nz = 256, the fused
    # vector-potential boundary kernel (conducting walls) with a uniform field
    monkeypatch.setenv("SX_XP", "10")
    P.case_mhd_substeps(emu_lib, tables, (64, 16, 256), ord=2, nsteps=1, impl=0, b0=(0.1, 0.0, 0.2))


def test_io_output_restart(emu_lib, tables, tmp_path):
    P.case_io_output_restart(emu_lib, tables, SMALL, tmp_path)


def test_wall_reconstructions(emu_lib, tables):
    P.case_wall_reconstructions(emu_lib, tables, SMALL)


def test_vacuum_walls(emu_lib, tables):
    P.case_vacuum_walls(emu_lib, tables, SMALL)


def test_more_diagnostics(emu_lib, tables):
    P.case_more_diagnostics(emu_lib, tables, SMALL)


@pytest.mark.parametrize("impl", [1, 0])
def test_rotbouss_substeps(emu_lib, tables, impl):
    P.case_rotbouss_substeps(emu_lib, tables, SMALL, ord=2, nsteps=1, impl=impl)


@pytest.mark.parametrize("impl", [1, 0])
def test_mhdbouss_substeps(emu_lib, tables, impl):
    P.case_mhdbouss_substeps(emu_lib, tables, SMALL, ord=2, nsteps=1, impl=impl)
    P.case_mhdbouss_substeps(emu_lib, tables, (16, 16, 64), ord=2, nsteps=1, bc=(0, 1), impl=impl)


def test_solver_output_restart(emu_lib, tables, tmp_path):
    P.case_solver_output_restart(emu_lib, tables, (16, 16, 64), tmp_path)


def test_golden_solvers(emu_lib, tables):
    P.case_golden_solvers(emu_lib, tables)


def test_full_size_properties_small(emu_lib, tables):
    # the property case of the GPU suite (512^3 there) on a grid the emulation finishes in seconds
    P.case_full_size_properties(emu_lib, tables, shape=(32, 16, 64), ord=2, dt=1e-3)


def test_boots_regridder(emu_lib, tables, tmp_path):
    # (nxt, nyt, nzt, ozt, nx, ny, nzp): A25-5 (period 52 -> 90), A50-5 with an odd period (153 -> 156),
    # periodic treatment with an odd old period (21 -> 42), identical grids
    P.case_boots(emu_lib, tables, [(16, 16, 27, 5, 32, 32, 46), (16, 16, 103, 5, 16, 32, 105), (16, 32, 21, 0, 32, 32, 41),
                                   (16, 16, 20, 0, 16, 16, 20)], tmp_path)
    import os
    P.case_boots_golden(emu_lib, tables, os.path.join(os.path.dirname(__file__), "golden", "boots_27_46.npz"))


def test_global_quantity_files(emu_lib, tables, tmp_path):
    assert P.case_global_files(emu_lib, tables, (16, 16, 64), tmp_path) == ["HD", "BOUSS", "MHDBOUSS"]


# The reference ships continuation tables for O = 3..9 matching points and C = 15..34 continuation points
# (tables/README.info); BASELINE.json uses A25-5 throughout.  Smallest, largest-C and largest-O members:
OTHER_TABLES = [(15, 3), (34, 8), (33, 9)]


@pytest.mark.parametrize("fc", OTHER_TABLES)
def test_other_fc_tables_operators(emu_lib, tables, fc):
    P.case_operators_other_table(emu_lib, tables, (16, 16, 64), *fc)


@pytest.mark.parametrize("fc", OTHER_TABLES)
@pytest.mark.parametrize("solver", ["hd", "bouss", "mhd"])
def test_other_fc_tables_substeps(emu_lib, tables, fc, solver):
    for impl in (0, 1):
        P.case_substeps_other_table(emu_lib, tables, (16, 16, 64), *fc, solver, impl=impl)


def test_other_fc_tables_long_pencils(emu_lib, tables):
    # the paired projection kernels (velocity: from nz = 256; vector potential, conducting walls) with d = 9 and d = 3
    for fc in ((33, 9), (15, 3)):
        for solver in ("hd", "mhd"):
            P.case_substeps_other_table(emu_lib, tables, (16, 16, 256), *fc, solver)


@pytest.mark.parametrize("solver", ["bouss", "mhd"])
def test_solver_diagnostics_several_steps(emu_lib, tables, solver):
    # bouss_global.f90 / mhd_global.f90 columns over a few steps (the GPU suite runs 100 steps at 64^3 against goldens)
    P.case_solver_diagnostics(emu_lib, tables, (16, 16, 64), solver, nsteps=2, every=1)


def test_adversarial_schedules_and_late_async_copies(tables):
    """The race check of the emulation (tests/emu/cuda_emu.h, SX_EMU_ADVERSARIAL): the threads of a block run in a random
    order between barriers, mbarrier waits really wait, and asynchronous copies complete as late as the program allows
    (bulk / tensor loads when a thread gets through the mbarrier wait, cp.async at the issuing thread's wait_group,
    tensor-map stores read their shared-memory source at wait_group.read), and the streams are lazy queues (the copy stream
    of sx_hd_step_host runs only through the events the compute stream waits for).  A missing barrier or wait changes the
    result; the bulk-copy kernels (forced on small grids, then the length-512 instantiations of the 512^3 bench under the
    default selection) and the host-buffer step must still agree with the oracle.  (The mode is read once per process: the
    cases run in a child.  tools/emu_racecheck.sh runs the whole emulation suite this way and shows that injected
    bugs -- a removed mbarrier wait, a removed wait_group.read -- are caught.)"""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import parity_cases as P\n"
            "from specter_b200 import api, build\n"
            "lib = api.Library(build.build_emu())\n"
            "P.case_hd_substeps(lib, %r, (16, 128, 128), ord=2, nsteps=1, impl=0)\n"
            "P.case_hd_substeps(lib, %r, (64, 16, 64), ord=2, nsteps=1, impl=1)\n"
            "P.case_mhd_substeps(lib, %r, (64, 16, 256), ord=2, nsteps=1, impl=0, b0=(0.1, 0.0, 0.2))\n"
            "P.case_hd_step_host(lib, %r, (32, 16, 64), pinned=True, nsteps=2, inflight=True)\n"
            "import os\n"
            "for k in ('SX_TMA_MIN', 'SX_XP', 'SX_PJ'): del os.environ[k]\n"
            "# default kernel selection: the template instantiations the 512^3 bench runs, one axis at a time\n"
            "for shape in ((512, 16, 64), (16, 512, 64), (16, 16, 512)):\n"
            "    P.case_hd_substeps(lib, %r, shape, ord=2, nsteps=1, impl=0)\n"
            "print('ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
                                tables, tables, tables, tables, tables)
    # 30 = random order + late asynchronous copies + lazy streams + eager copy stream: 14 = (operations of a stream run only when the host, or an
    # event another stream waits for, needs them: work that no event orders before its consumer has not run by then)
    # + 16: the copy stream of the host-buffer step runs as EARLY as its event waits allow while the compute stream is lazy
    env = dict(os.environ, SX_EMU_ADVERSARIAL="30", SX_TMA_MIN="16", SX_XP="10", SX_PJ="10")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
