"""Randomised sweep on the CPU-thread emulation: shapes x solvers x fused / per-operator x tuning-knob sets, one RK2 step each against the
oracle at the parity tolerances.   python tools/emu_sweep.py [seed] [count]   (SX_EMU_ADVERSARIAL applies; round 2: 150 + 150 combinations, 0 failures)"""
import sys, os, time, itertools, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import parity_cases as P
from specter_b200 import api, build
lib = api.Library(build.build_emu())
T = os.path.join(ROOT, 'tests', 'golden', 'tables')
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
shapes = [(16,16,64),(32,16,64),(16,32,64),(64,16,64),(16,64,64),(16,16,128),(128,16,64),(16,128,64),(32,32,128),(16,16,256),(256,16,64),(16,256,64)]
knobsets = [{}, {"SX_TMA_MIN":"16"}, {"SX_XP":"10"}, {"SX_PJ":"10"}, {"SX_XP":"9","SX_PJ":"9"}, {"SX_TMA_MIN":"16","SX_INV_STAGES":"1"}, {"SX_TMA_MIN":"16","SX_INV_STAGES":"2"},
            {"SX_TILE_PF":"0"}, {"SX_TILE_PF":"15"}, {"SX_TMA":"0"}, {"SX_TMA":"7","SX_TMA_MIN":"16"}]
fails = 0; n = 0
t0=time.time()
combos = [(s,k,sol,impl,ordr) for s in shapes for k in knobsets for sol in ("hd","bouss","mhd") for impl in (0,1) for ordr in (2,)]
random.shuffle(combos)
for shape, knobs, solver, impl, ordr in combos[:int(sys.argv[2]) if len(sys.argv)>2 else 120]:
    for k in ("SX_TMA_MIN","SX_XP","SX_PJ","SX_INV_STAGES","SX_TILE_PF","SX_TMA"): os.environ.pop(k, None)
    os.environ.update(knobs)
    n += 1
    try:
        {"hd": P.case_hd_substeps, "bouss": P.case_bouss_substeps, "mhd": P.case_mhd_substeps}[solver](lib, T, shape, ord=ordr, nsteps=1, impl=impl)
    except api.SpecterError as e:
        print("REFUSED", shape, knobs, solver, impl, str(e)[:100], flush=True)
    except AssertionError as e:
        fails += 1
        print("FAIL", shape, knobs, solver, impl, str(e)[:100], flush=True)
print(f"{n} combos, {fails} failures, {time.time()-t0:.0f}s")
