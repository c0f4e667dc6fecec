#!/usr/bin/env python
"""One 256-length case through the bulk-copy / tensor-map kernels (knob_tma_min = 256) for compute-sanitizer:
HD, BOUSS and MHD fused RK2 substeps on 256x32x256 against nothing but themselves (finite results, launches > 0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from specter_b200 import api  # noqa: E402

for solver in ("hd", "bouss", "mhd"):
    p = api.Plan(256, 32, 256, 25, 5, ord=2, tdir=bench.TABLES, device=0)
    bench.device_state(p, solver)
    n0 = p.launch_count
    if solver == "hd":
        p.hd_step(1e-3, 1e-3)
        st = p.hd_get_state()
    elif solver == "bouss":
        p.bouss_step(1e-3, 1e-3, 1e-3)
        st = p.bouss_get_state()
    else:
        p.mhd_step(1e-3, 1e-3, 5e-3)
        st = p.mhd_get_state()
    assert all(np.isfinite(a).all() for a in st)
    print(solver, "ok", p.launch_count - n0, "launches")
    p.close()
