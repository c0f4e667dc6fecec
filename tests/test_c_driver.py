"""The C ABI from a compiled host language: examples/hd_driver.c (plain C99 against include/specter_b200.h, the loop a
maintainer's Fortran driver runs through the ISO_C_BINDING module of INTEGRATION.md) continues a run from the
reference-format field files, steps it, prints the global quantities and writes the BIN block; every output is
compared with the oracle doing the same.  CPU: linked against the kernel emulation build; `-m gpu`: against the
nvcc-built library."""
import os
import subprocess

import numpy as np
import pytest

from oracle import specter_oracle as O
from specter_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_FIELD, TOL_DIAG = 1e-11, 1e-9


def run_driver(libpath, tables, tmp_path, shape=(16, 16, 64), ord=2, nsteps=3, dt=1e-3, nu=1e-3, f0=1.0, env=None):
    nx, ny, nz = shape
    libdir, libname = os.path.dirname(libpath), os.path.basename(libpath)[3:-3]
    exe = str(tmp_path / "hd_driver")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "examples", "hd_driver.c"), "-L", libdir, "-l" + libname,
                         "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    idir, odir, rdir = tmp_path / "in", tmp_path / "out", tmp_path / "ref"
    for d in (idir, odir, rdir):
        d.mkdir()
    # a previous run's files, written in the reference's format by the oracle
    g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=ord)
    s0 = O.make_hd_state(g)
    O.hd_step(g, s0, dt, nu)                        # so that pr is not identically zero in the files
    O.hd_output(g, s0, str(idir), "0001", dt)
    r = subprocess.run([exe, tables, str(idir), str(odir), str(nx), str(ny), str(nz), "25", "5", str(ord), "1.0", "0.5", "1.0",
                        repr(dt), repr(nu), repr(f0), str(nsteps), "1", "0001", "0002"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    # the oracle through the same sequence: restart, forcing, steps with the global quantities, output
    vx, vy, vz, pr = O.hd_restart(g, str(idir), "0001", dt)
    fx, fy, fz = O.initialfv(g, f0)
    s = O.HDState(vx, vy, vz, pr, fx, fy, fz)
    ref_rows = []
    for t in range(nsteps):
        ref_rows.append(O.hdcheck(g, s.vx, s.vy, s.vz, s.fx, s.fy, s.fz) + O.vdiagnostic(g, s.vx, s.vy, s.vz))
        O.hd_step(g, s, dt, nu)
    O.hd_output(g, s, str(rdir), "0002", dt)
    rows = [ln.replace("|", " ").split() for ln in r.stdout.strip().splitlines()]
    assert len(rows) == nsteps
    for t, (row, ref) in enumerate(zip(rows, ref_rows)):
        got = np.array([float(x) for x in row])
        assert abs(got[0] - t * dt) < 1e-12
        assert np.allclose(got[1:4], ref[:3], rtol=TOL_DIAG, atol=0)                  # energy, dissipation, injection
        assert np.abs(got[4:] - np.array(ref[3:])).max() <= TOL_DIAG * abs(ref[0])     # residuals on the scale of the energy
    nph = nz - 25
    for name in ("vx", "vy", "vz", "pr"):
        a = np.fromfile(O.io_path(str(odir), name, "0002"))
        b = np.fromfile(O.io_path(str(rdir), name, "0002"))
        assert a.size == b.size == nx * ny * nph
        err = np.abs(a - b).max() / np.abs(b).max()
        assert err < (100 * TOL_FIELD if name == "pr" else TOL_FIELD), (name, err)    # p = p'/dt, see parity_cases
    bench = open(odir / "benchmark.txt").read().split("\n")
    assert len([ln for ln in bench if ln.strip()]) == 2 and bench[1].split()[:4] == [str(nx), str(ny), str(nz), str(nsteps)]
    return r.stderr


def test_c_driver_on_emulated_kernels(tables, tmp_path):
    run_driver(build.build_emu(), tables, tmp_path)


@pytest.mark.gpu
def test_c_driver_on_gpu(cuda_lib, tables, tmp_path):
    log = run_driver(api.LIB_PATH, tables, tmp_path, shape=(64, 64, 64), nsteps=3)
    assert "kernel launches" in log


# ---- examples/solver_driver.c: the same loop for every solver of the reference --------------------------------------
def run_solver_driver(libpath, tables, tmp_path, solver, shape=(16, 16, 64), nsteps=2, dt=1e-3, nu=1e-3, kappa=1e-3, mu=5e-3,
                      f0=1.0, bc=(0, 0), b0=(0.0, 0.0, 0.1), omega=(0.3, -0.2, 1.5)):
    import parity_cases as P
    nx, ny, nz = shape
    libdir, libname = os.path.dirname(libpath), os.path.basename(libpath)[3:-3]
    exe = str(tmp_path / "solver_driver")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "examples", "solver_driver.c"), "-L", libdir, "-l" + libname,
                         "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    idir, odir, rdir = tmp_path / ("in_" + solver), tmp_path / ("out_" + solver), tmp_path / ("ref_" + solver)
    for d in (idir, odir, rdir):
        d.mkdir()
    g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=2)
    g.load_neumann()
    sca, mag = solver in ("BOUSS", "ROTBOUSS", "MHDBOUSS"), solver in ("MHD", "MHDBOUSS")
    zero = lambda: np.zeros(g.cshape(), dtype=np.complex128)

    def step(s):
        if solver == "HD":
            O.hd_step(g, s, dt, nu)
        elif solver == "BOUSS":
            O.bouss_step(g, s, dt, nu, kappa)
        elif solver == "ROTBOUSS":
            O.rotbouss_step(g, s, dt, nu, kappa, omega=omega)
        elif solver == "MHD":
            O.mhd_step(g, s, dt, nu, mu, b0=b0)
        else:
            O.mhdbouss_step(g, s, dt, nu, mu, kappa, b0=b0, bczsta=bc[0], bczend=bc[1])

    # a previous run's files in the reference's format (one oracle step so that pr and ph are not identically zero)
    s0 = {"HD": O.make_hd_state, "BOUSS": O.make_bouss_state, "ROTBOUSS": O.make_bouss_state, "MHD": O.make_mhd_state,
          "MHDBOUSS": O.make_mhdbouss_state}[solver](g)
    step(s0)
    O.solver_output(g, s0, str(idir), "0001", dt)
    kinds = [P.B_KIND[bc[0]], P.B_KIND[bc[1]]]
    args = [exe, solver, tables, str(idir), str(odir), str(nx), str(ny), str(nz), "25", "5", "2", "1.0", "0.5", "1.0", repr(dt),
            repr(nu), repr(kappa), repr(mu), repr(f0), str(nsteps), "1", "0001", "0002"] + kinds + [repr(x) for x in b0 + omega]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the oracle through the same sequence
    w = O.solver_restart(g, str(idir), "0001", dt, scalar=sca, magnetic=mag)
    fx, fy, fz = O.initialfv(g, f0)
    if solver == "HD":
        s = O.HDState(w["vx"], w["vy"], w["vz"], w["pr"], fx, fy, fz)
    elif not mag:
        s = O.BoussState(w["vx"], w["vy"], w["vz"], w["pr"], fx, fy, fz, w["th"], zero())
    elif solver == "MHD":
        s = O.MhdState(w["vx"], w["vy"], w["vz"], w["pr"], fx, fy, fz, w["ax"], w["ay"], w["az"], w["ph"], zero(), zero(), zero())
    else:
        s = O.MhdBoussState(w["vx"], w["vy"], w["vz"], w["pr"], fx, fy, fz, w["ax"], w["ay"], w["az"], w["ph"], zero(), zero(),
                            zero(), w["th"], zero())
    eng = O.energy(g, s.vx, s.vy, s.vz, 1)
    for t in range(1, nsteps + 1):
        O.solver_global(g, s, solver, str(rdir), t, dt, *bc)
        step(s)
    O.solver_output(g, s, str(rdir), "0002", dt)
    P.compare_global_dirs(odir, rdir, solver, nsteps, eng)
    nph = nz - 25
    outs = sorted(n for n in os.listdir(rdir) if n.endswith(".out"))
    assert outs == sorted(n for n in os.listdir(odir) if n.endswith(".out"))
    for fn in outs:
        name = fn.split(".")[0]
        a, b = np.fromfile(str(odir / fn)), np.fromfile(str(rdir / fn))
        assert a.size == b.size == nx * ny * nph
        err = np.abs(a - b).max() / np.abs(b).max()
        tol = 100 * TOL_FIELD if name in ("pr", "ph") else (P.TOL_RECONTINUED if name == "th" else 10 * TOL_FIELD)
        assert err < tol, (solver, name, err)
    return r.stderr


@pytest.mark.parametrize("solver,bc", [("ROTBOUSS", (0, 0)), ("MHDBOUSS", (0, 1))])
def test_solver_driver_on_emulated_kernels(tables, tmp_path, solver, bc):
    # the two that cover every branch of the driver between them (BOUSS and MHD run in the GPU suite); one step: the cost
    # here is the emulated diagnostics
    run_solver_driver(build.build_emu(), tables, tmp_path, solver, nsteps=1, bc=bc)
