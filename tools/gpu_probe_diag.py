"""Few-second GPU probe: BOUSS / MHD diagnostics over 100 steps against the committed goldens (tests/golden/solvers64_diag100.json)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import parity_cases as P  # noqa: E402
from specter_b200 import api  # noqa: E402

lib = api.load_library()
T = os.path.join(ROOT, "tests", "golden", "tables")
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "solvers64_diag100.json")))
for solver in ("bouss", "mhd"):
    t = time.time()
    try:
        _, _, worst = P.case_solver_diagnostics(lib, T, (64, 64, 64), solver, nsteps=100, every=10, golden=gold[solver]["rows"])
        print(f"ok   {solver} {time.time() - t:.1f}s worst {max(worst.values()):.1e} ({max(worst, key=worst.get)})", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"FAIL {solver} {type(e).__name__}: {str(e)[:200]}", flush=True)
