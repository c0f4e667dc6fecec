"""Generates the committed golden fixtures from the oracle (the reference itself cannot be built in
this image -- no Fortran/MPI/FFTW -- and ships no golden vectors; see DESIGN.md "Oracle").

    python tests/golden/make_golden.py

hd64_diag100.json   config 1 (HD 64^3, C=25 d=5, RK2, dt=1e-3, nu=1e-3, Lx=1 Ly=.5 Lz=1, seed=1000):
                    balance.txt columns (<v^2>, <w^2>-like, <v.f>) and <(div v)^2> every 10 steps.
hd64_step1.npz      the same run, spectral fields after the first time step on a coarse sub-sample.
solvers32_step1.npz one RK2 step of the other solvers on 32x32x64.
boots_27_46.npz     the BOOTS regridder (tools/boots.fpp): a seeded field on 16x16x27 -> 32x32x46 (A25-5 continuation) and on
                    16x32x21 -> 32x32x41 (periodic treatment, odd old period), outputs sub-sampled.
                    `python tests/golden/make_golden.py boots` regenerates this file only.
solvers64_diag100.json  BOUSS and MHD (conducting walls) on 64^3, 100 RK2 steps, the columns of bouss_global.f90 /
                    mhd_global.f90 every 10 steps (`python tests/golden/make_golden.py diag` regenerates this file only).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import specter_oracle as O  # noqa: E402



sys.path.insert(0, os.path.dirname(HERE))
from parity_cases import boots_golden_inputs as boots_inputs  # noqa: E402  (the seeded old-grid fields, shared with the tests)


def boots_goldens():
    a, b = boots_inputs()
    tab = os.path.join(HERE, "tables")
    ra = O.boots_regrid(a, 32, 32, 46, 5, tab)
    rb = O.boots_regrid(b, 32, 32, 41, 0, tab)
    np.savez_compressed(os.path.join(HERE, "boots_27_46.npz"), a=ra[::3, ::4, ::4], b=rb[::3, ::4, ::4])
    print("boots_27_46.npz:", ra.shape, rb.shape)


def solver_diag_goldens():
    import parity_cases as P
    out = {"config": "64^3 Cz=25 oz=5 RK2 dt=1e-3 nu=1e-3 kappa=1e-3 mu=5e-3 Lx=1 Ly=0.5 Lz=1 seed=1000, sampled every 10 steps"}
    for solver in ("bouss", "mhd"):
        gd = O.Grid(64, 64, 64, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=os.path.join(HERE, "tables"), ord=2)
        rows = P.oracle_solver_diagnostics(gd, solver, nsteps=100, every=10)
        out[solver] = {"columns": ["step"] + list(P._DIAG_COLS[solver]), "rows": [[float(x) for x in r] for r in rows]}
        print(solver, rows[-1])
    with open(os.path.join(HERE, "solvers64_diag100.json"), "w") as f:
        json.dump(out, f, indent=1)


if len(sys.argv) > 1 and sys.argv[1] == "boots":
    boots_goldens()
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "diag":
    solver_diag_goldens()
    sys.exit(0)

g = O.Grid(64, 64, 64, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=os.path.join(HERE, "tables"), ord=2)
s = O.make_hd_state(g)
rows = []
for t in range(100):
    O.hd_step(g, s, 1e-3, 1e-3)
    if t == 0:
        np.savez_compressed(os.path.join(HERE, "hd64_step1.npz"), vx=s.vx[::4, ::4, ::2], vy=s.vy[::4, ::4, ::2],
                            vz=s.vz[::4, ::4, ::2], pr=s.pr[::4, ::4, ::2])
    if (t + 1) % 10 == 0:
        rows.append([t + 1, *O.hdcheck(g, s.vx, s.vy, s.vz, s.fx, s.fy, s.fz), O.divergence(g, s.vx, s.vy, s.vz)])
        print(rows[-1])
with open(os.path.join(HERE, "hd64_diag100.json"), "w") as f:
    json.dump({"config": "HD 64^3 Cz=25 oz=5 RK2 dt=1e-3 nu=1e-3 Lx=1 Ly=0.5 Lz=1 seed=1000 f0=1",
               "columns": ["step", "energy", "enstrophy_ref_quirk", "injection", "divergence"], "rows": rows}, f, indent=1)


# ---- solvers32_step1.npz: one RK2 step of every other solver on 32x32x64 (same box, seed, dt, nu; kappa=1e-3, mu=5e-3,
# omega=(0.3,-0.2,1.5), b0=(0,0,0.1)), spectral fields sub-sampled, plus the global-output columns of that state
def solver_goldens():
    gs = O.Grid(32, 32, 64, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=os.path.join(HERE, "tables"), ord=2)
    gs.load_neumann()
    out = {}
    sub = (slice(None, None, 4), slice(None, None, 4), slice(None, None, 4))

    def put(tag, s, names):
        for n in names:
            out[f"{tag}_{n}"] = getattr(s, n)[sub]

    b = O.make_bouss_state(gs)
    O.bouss_step(gs, b, 1e-3, 1e-3, 1e-3)
    put("bouss", b, ("vx", "vy", "vz", "th"))
    out["bouss_pscheck"] = np.array(O.pscheck(gs, b.th, b.vz))
    r = O.make_bouss_state(gs)
    O.rotbouss_step(gs, r, 1e-3, 1e-3, 1e-3, omega=(0.3, -0.2, 1.5))
    put("rotbouss", r, ("vx", "vy", "vz", "th"))
    for tag, bc in (("mhd", (0, 0)), ("mhdvac", (1, 1))):
        m = O.make_mhdbouss_state(gs)
        O.mhdbouss_step(gs, m, 1e-3, 1e-3, 5e-3, 1e-3, b0=(0.0, 0.0, 0.1), bczsta=bc[0], bczend=bc[1])
        put(tag + "bouss", m, ("vx", "vy", "vz", "ax", "ay", "az", "th"))
        out[tag + "bouss_mhdcheck"] = np.array(O.mhdcheck(gs, m.vx, m.vy, m.vz, m.ax, m.ay, m.az))
        d = O.bdiagnostic(gs, m.ax, m.ay, m.az, *bc)
        out[tag + "bouss_bdiag"] = np.array(d["conducting" if bc == (0, 0) else "vacuum"])
    m = O.make_mhd_state(gs)
    O.mhd_step(gs, m, 1e-3, 1e-3, 5e-3)
    put("mhd", m, ("vx", "vy", "vz", "ax", "ay", "az"))
    np.savez_compressed(os.path.join(HERE, "solvers32_step1.npz"), **out)
    print("solvers32_step1.npz:", sorted(out))


solver_goldens()
boots_goldens()
solver_diag_goldens()
