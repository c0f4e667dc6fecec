"""GPU parity proper: the nvcc-built sm_100a library through the C ABI against the oracle."""
import json
import os

import numpy as np
import pytest

import parity_cases as P

pytestmark = pytest.mark.gpu
CFG1 = (64, 64, 64)  # BASELINE.json configs[0]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fft1d_z(cuda_lib, tables):
    for shape in (CFG1, (32, 16, 128), (16, 32, 256), (16, 16, 512), (16, 16, 1024)):
        P.case_fft1d_z(cuda_lib, tables, shape)
    P.case_fft1d_z(cuda_lib, tables, (16, 16, 32), Cz=0)


def test_fft3d(cuda_lib, tables):
    for shape in (CFG1, (128, 32, 64), (32, 256, 64), (512, 16, 64), (16, 512, 128)):
        P.case_fft3d(cuda_lib, tables, shape)
    P.case_fft3d(cuda_lib, tables, (64, 32, 32), Cz=0)


def test_fft_known_answer(cuda_lib, tables):
    P.case_fft_known_answer(cuda_lib, tables, 64)


def test_spectral_ops(cuda_lib, tables):
    P.case_spectral_ops(cuda_lib, tables, CFG1)


def test_nonlinear(cuda_lib, tables):
    P.case_nonlinear(cuda_lib, tables, CFG1)
    P.case_nonlinear(cuda_lib, tables, (128, 64, 128))


def test_projection(cuda_lib, tables):
    P.case_projection(cuda_lib, tables, CFG1)


def test_diagnostics(cuda_lib, tables):
    P.case_diagnostics(cuda_lib, tables, CFG1)


@pytest.mark.parametrize("impl", [1, 0])
def test_hd_substeps_cfg1(cuda_lib, tables, impl):
    P.case_hd_substeps(cuda_lib, tables, CFG1, ord=2, nsteps=3, impl=impl)


def test_hd_substeps_rk4(cuda_lib, tables):
    P.case_hd_substeps(cuda_lib, tables, (128, 128, 128), ord=4, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (64, 32, 256), ord=4, nsteps=1, impl=0, walls=((0.2, -0.1), (-0.3, 0.1)))


def test_hd_step_host(cuda_lib, tables):
    P.case_hd_step_host(cuda_lib, tables, CFG1)


def test_hd_diagnostics_100_steps_cfg1(cuda_lib, tables):
    """Config 1, 100 RK2 steps: balance.txt / noslip_diagnostic columns vs the committed oracle goldens."""
    with open(os.path.join(GOLD, "hd64_diag100.json")) as f:
        gold = json.load(f)
    rows = P.case_hd_diagnostics_100(cuda_lib, tables, CFG1, nsteps=100, ord=2, impl=0, golden=gold["rows"])
    assert rows.shape[0] == 10


def test_advect_vector(cuda_lib, tables):
    P.case_advect_vector(cuda_lib, tables, CFG1)


def test_scalar_vecpot_bc(cuda_lib, tables):
    P.case_scalar_vecpot_bc(cuda_lib, tables, CFG1)


@pytest.mark.parametrize("impl", [1, 0])
def test_bouss_substeps_cfg1(cuda_lib, tables, impl):
    P.case_bouss_substeps(cuda_lib, tables, CFG1, ord=2, nsteps=2, impl=impl)


def test_bouss_substeps_rk4(cuda_lib, tables):
    P.case_bouss_substeps(cuda_lib, tables, (128, 64, 128), ord=4, nsteps=1, impl=0)


@pytest.mark.parametrize("impl", [1, 0])
def test_mhd_substeps_cfg1(cuda_lib, tables, impl):
    P.case_mhd_substeps(cuda_lib, tables, CFG1, ord=2, nsteps=2, impl=impl)


def test_mhd_substeps_rk4_uniform_field(cuda_lib, tables):
    P.case_mhd_substeps(cuda_lib, tables, (64, 128, 128), ord=4, nsteps=1, impl=0, b0=(0.1, 0.0, 0.2))


def test_hd_substeps_length_512_kernels(cuda_lib, tables):
    """The kernel instantiations the 512^3 bench runs (length-512 transforms along each axis in turn, bulk-copy
    x pass from nx = 256), on grids small enough for the oracle."""
    P.case_hd_substeps(cuda_lib, tables, (512, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (16, 512, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (16, 16, 512), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (256, 32, 128), ord=4, nsteps=1, impl=0)


def test_bouss_mhd_substeps_long_x(cuda_lib, tables):
    P.case_bouss_substeps(cuda_lib, tables, (256, 32, 64), ord=2, nsteps=1, impl=0)
    P.case_bouss_substeps(cuda_lib, tables, (512, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_mhd_substeps(cuda_lib, tables, (256, 32, 64), ord=2, nsteps=1, impl=0)


def test_hd_substeps_length_1024_2048_kernels(cuda_lib, tables):
    """Transform lengths of the multi-GPU configurations (1024x1024x512, 2048x2048x1024) along each axis in turn."""
    P.case_hd_substeps(cuda_lib, tables, (1024, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (2048, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (16, 1024, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (16, 2048, 64), ord=2, nsteps=1, impl=0)
    P.case_hd_substeps(cuda_lib, tables, (16, 16, 1024), ord=2, nsteps=1, impl=0)


def test_mhd_bouss_baseline_lengths(cuda_lib, tables):
    """The kernel instantiations of BASELINE configs[3] (MHD 512^3: cross-product x pass, conducting-wall kernels at
    nz = 512) and configs[2] (BOUSS 1024x1024x512: four-component x pass at nx = 1024), one axis at a time on grids the
    oracle finishes in seconds."""
    P.case_mhd_substeps(cuda_lib, tables, (512, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_mhd_substeps(cuda_lib, tables, (16, 512, 64), ord=2, nsteps=1, impl=0)
    P.case_mhd_substeps(cuda_lib, tables, (16, 16, 512), ord=2, nsteps=1, impl=0)
    P.case_bouss_substeps(cuda_lib, tables, (1024, 16, 64), ord=2, nsteps=1, impl=0)
    P.case_bouss_substeps(cuda_lib, tables, (16, 1024, 64), ord=2, nsteps=1, impl=0)
    P.case_bouss_substeps(cuda_lib, tables, (16, 16, 512), ord=2, nsteps=1, impl=0)


def test_io_output_restart(cuda_lib, tables, tmp_path):
    P.case_io_output_restart(cuda_lib, tables, CFG1, tmp_path)


def test_wall_reconstructions(cuda_lib, tables):
    P.case_wall_reconstructions(cuda_lib, tables, CFG1)


def test_vacuum_walls(cuda_lib, tables):
    P.case_vacuum_walls(cuda_lib, tables, CFG1)


def test_more_diagnostics(cuda_lib, tables):
    P.case_more_diagnostics(cuda_lib, tables, CFG1)


@pytest.mark.parametrize("impl", [1, 0])
def test_rotbouss_substeps_cfg1(cuda_lib, tables, impl):
    P.case_rotbouss_substeps(cuda_lib, tables, CFG1, ord=2, nsteps=2, impl=impl)


def test_rotbouss_substeps_rk4_moving_walls(cuda_lib, tables):
    P.case_rotbouss_substeps(cuda_lib, tables, (128, 64, 128), ord=4, nsteps=1, impl=0, walls=((0.2, -0.1), (-0.3, 0.1)))


@pytest.mark.parametrize("impl", [1, 0])
def test_mhdbouss_substeps_cfg1(cuda_lib, tables, impl):
    P.case_mhdbouss_substeps(cuda_lib, tables, CFG1, ord=2, nsteps=2, impl=impl)
    P.case_mhdbouss_substeps(cuda_lib, tables, (32, 32, 64), ord=2, nsteps=1, bc=(1, 1), impl=impl)


def test_mhdbouss_substeps_long_x(cuda_lib, tables):
    # the bulk-copy scalar-advection x pass (NC = 1, from nx = 256) next to the cross-product passes
    P.case_mhdbouss_substeps(cuda_lib, tables, (256, 32, 64), ord=2, nsteps=1, impl=0)


def test_solver_output_restart(cuda_lib, tables, tmp_path):
    P.case_solver_output_restart(cuda_lib, tables, (32, 32, 64), tmp_path)


def test_golden_fixtures(cuda_lib, tables):
    P.case_golden_solvers(cuda_lib, tables)
    P.case_golden_hd_step1(cuda_lib, tables)


def test_full_size_properties_hd512(cuda_lib, tables):
    """BASELINE.json configs[1] (HD 512^3 RK4) through size-independent properties."""
    P.case_full_size_properties(cuda_lib, tables)


def test_boots_regridder(cuda_lib, tables, tmp_path):
    # (nxt, nyt, nzt, ozt, nx, ny, nzp); tile edges of the dense z operator in every direction
    P.case_boots(cuda_lib, tables, [(16, 16, 27, 5, 32, 32, 46), (64, 32, 79, 5, 128, 64, 136), (32, 16, 103, 5, 32, 64, 105),
                                    (16, 32, 21, 0, 32, 32, 41), (128, 128, 131, 5, 256, 128, 256)], tmp_path)


def test_goto_domain(cuda_lib, tables):
    P.case_goto_domain(cuda_lib, tables, CFG1)
    P.case_goto_domain(cuda_lib, tables, (32, 16, 512))


def test_global_quantity_files(cuda_lib, tables, tmp_path):
    assert P.case_global_files(cuda_lib, tables, (32, 32, 64), tmp_path) == ["HD", "BOUSS", "MHDBOUSS"]


def test_normalisations(cuda_lib, tables):
    P.case_normalisations(cuda_lib, tables, CFG1)


@pytest.mark.parametrize("solver", ["bouss", "mhd"])
def test_solver_diagnostics_100_steps_cfg1(cuda_lib, tables, solver):
    """BOUSS and MHD (conducting walls) on 64^3, 100 RK2 steps: the columns of bouss_global.f90 / mhd_global.f90 every
    10 steps against the committed oracle goldens (tests/golden/solvers64_diag100.json), 1e-9."""
    with open(os.path.join(GOLD, "solvers64_diag100.json")) as f:
        gold = json.load(f)
    rows, _, _ = P.case_solver_diagnostics(cuda_lib, tables, CFG1, solver, nsteps=100, every=10, golden=gold[solver]["rows"])
    assert rows.shape[0] == 10


# ---- the reference's other continuation tables (O = 3..9, C = 15..34; tables/README.info) -------------------------
@pytest.mark.parametrize("fc", [(15, 3), (34, 8), (33, 9)])
def test_other_fc_tables(cuda_lib, tables, fc):
    P.case_operators_other_table(cuda_lib, tables, CFG1, *fc)
    for solver, impl in (("hd", 0), ("hd", 1), ("bouss", 0), ("mhd", 0)):
        P.case_substeps_other_table(cuda_lib, tables, CFG1, *fc, solver, impl=impl, draws=2)
    if fc == (34, 8):
        return
    # the long-line kernel families (bulk-copy tiles, x-pass ring, paired projections), one axis at a time
    for shape in ((16, 16, 512), (512, 16, 64), (16, 512, 64)):
        for solver in ("hd", "mhd"):
            P.case_substeps_other_table(cuda_lib, tables, shape, *fc, solver, draws=2)


# ---- examples/solver_driver.c on the B200 (kept in this file, after the cases above: the suite runs with -x) ---------
@pytest.mark.parametrize("solver,bc", [("HD", (0, 0)), ("BOUSS", (0, 0)), ("ROTBOUSS", (0, 0)), ("MHD", (0, 0)), ("MHDBOUSS", (1, 1))])
def test_solver_driver_on_gpu(cuda_lib, tables, tmp_path, solver, bc):
    from specter_b200 import api
    from test_c_driver import run_solver_driver
    log = run_solver_driver(api.LIB_PATH, tables, tmp_path, solver, shape=(32, 32, 64), nsteps=2, bc=bc)
    assert "kernel launches" in log
