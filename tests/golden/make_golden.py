"""Generates the committed golden fixtures from the oracle (the reference itself cannot be built in
this image -- no Fortran/MPI/FFTW -- and ships no golden vectors; see DESIGN.md "Oracle").

    python tests/golden/make_golden.py

hd64_diag100.json   config 1 (HD 64^3, C=25 d=5, RK2, dt=1e-3, nu=1e-3, Lx=1 Ly=.5 Lz=1, seed=1000):
                    balance.txt columns (<v^2>, <w^2>-like, <v.f>) and <(div v)^2> every 10 steps.
hd64_step1.npz      the same run, spectral fields after the first time step on a coarse sub-sample.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import specter_oracle as O  # noqa: E402

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
