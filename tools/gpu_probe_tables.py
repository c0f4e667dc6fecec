"""40-second GPU probe of the cases this session added for the reference's other continuation tables (the riskiest of the
new -m gpu cases: kernels with d = 9 / C = 33 have never run on hardware).  Prints one line per case."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
t0 = time.time()
import parity_cases as P  # noqa: E402
from specter_b200 import api  # noqa: E402

lib = api.load_library()
T = os.path.join(ROOT, "tests", "golden", "tables")
print(f"loaded {time.time() - t0:.1f}s", flush=True)
for name, fn in (("hd 64^3 A33-9", lambda: P.case_substeps_other_table(lib, T, (64, 64, 64), 33, 9, "hd", draws=1)),
                 ("hd 16x16x512 A33-9", lambda: P.case_substeps_other_table(lib, T, (16, 16, 512), 33, 9, "hd", draws=1)),
                 ("mhd 64^3 A33-9", lambda: P.case_substeps_other_table(lib, T, (64, 64, 64), 33, 9, "mhd", draws=1)),
                 ("operators 64^3 A34-8", lambda: P.case_operators_other_table(lib, T, (64, 64, 64), 34, 8)),
                 ("hd 512x16x64 A15-3", lambda: P.case_substeps_other_table(lib, T, (512, 16, 64), 15, 3, "hd", draws=1)),
                 ("mhd 16x16x512 A33-9", lambda: P.case_substeps_other_table(lib, T, (16, 16, 512), 33, 9, "mhd", draws=1))):
    t = time.time()
    try:
        fn()
        print(f"ok   {name} {time.time() - t:.1f}s", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"FAIL {name} {type(e).__name__}: {str(e)[:200]}", flush=True)
print(f"total {time.time() - t0:.1f}s", flush=True)
