"""The C-ABI library loads without a GPU and exports every symbol include/specter_b200.h declares;
the ctypes table of the host-side mirror agrees with the header; plan creation fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from specter_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "specter_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_bindings_agree():
    names = header_functions()
    assert len(names) > 40
    assert sorted(api.SIGNATURES) == names


def test_cuda_library_exports_every_symbol():
    lib = build.build_cuda()
    api._preload_nccl()   # the same order api.Library uses: torch's NCCL first, so a later `import torch` still resolves
    dll = ctypes.CDLL(lib)
    for name in header_functions():
        assert hasattr(dll, name), name
    dll.sx_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.sx_version()


def test_header_is_plain_c(tmp_path):
    # the boundary is a C ABI: the header must compile as C99 on its own (no C++, no CUDA, no torch types), which is
    # what a cgo / ISO_C_BINDING / ctypes user sees
    src = tmp_path / "use_header.c"
    src.write_text('#include "specter_b200.h"\n'
                   'int use(void) { sx_config c; sx_plan* p = 0; c.nx = 16; (void)p; return (int)sizeof c + c.nx; }\n')
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "use_header.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_fortran_module_is_current_and_complete():
    # include/specter_b200_mod.f90 is generated from the header (tools/gen_fortran_module.py): up to date, one
    # BIND(C) interface per prototype, same number of dummies as the ctypes table has argument types
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_module.py"), "--check"])
    assert r.returncode == 0, "run python tools/gen_fortran_module.py"
    raw = open(os.path.join(ROOT, "include", "specter_b200_mod.f90")).read()
    text = raw.replace("&\n", " ")
    found = dict(re.findall(r"FUNCTION (sx_[a-z0-9_]+)\(([^)]*)\)\s+BIND\(C,\s+NAME='\1'\)", text))
    assert sorted(found) == header_functions()
    for name, dummies in found.items():
        n = len([d for d in dummies.split(",") if d.strip()])
        assert n == len(api.SIGNATURES[name]), name
    assert max(len(ln) for ln in raw.splitlines()) <= 132      # free-form source line limit


def test_no_cpu_fallback(tables):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.SpecterError):
        api.Plan(16, 16, 64, 25, 5, tdir=tables)


def test_product_loader_refuses_the_emulation_build(emu_lib):
    # the tests hand the emulation build to api.Library explicitly; the product entry point (load_library) must not
    # accept it, whatever SPECTER_B200_LIB says
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\nfrom specter_b200 import api\n"
            "try:\n    api.load_library()\nexcept api.SpecterError as e:\n    print('refused:', e)\n" % ROOT)
    env = dict(os.environ, SPECTER_B200_LIB=emu_lib.path)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "refused:" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_argument_errors_on_emulated_build(emu_lib, tables):
    with pytest.raises(api.SpecterError, match="power"):
        api.Plan(24, 16, 64, 25, 5, tdir=tables, lib=emu_lib)
    with pytest.raises(api.SpecterError, match="table"):
        api.Plan(16, 16, 64, 25, 5, tdir="/nonexistent", lib=emu_lib)
    with pytest.raises(api.SpecterError, match="Mismatch"):
        api.Plan(16, 16, 64, 25, 0, tdir=tables, lib=emu_lib)
    with pytest.raises(api.SpecterError, match="too many ranks"):
        api.Plan(16, 16, 64, 25, 5, tdir=tables, nprocs=12, myrank=11, lib=emu_lib)
    p = api.Plan(16, 16, 64, 25, 5, tdir=tables, lib=emu_lib)
    a = p.spectral()
    with pytest.raises(api.SpecterError, match="dir"):
        p.derivk(a, a, 4)
    with pytest.raises(api.SpecterError, match="Unsupported BC"):
        p.sol_project(a, a, a, a, 1, 1, 1)
    with pytest.raises(api.SpecterError, match="substep"):
        p.hd_rkstep2(3, 1e-3, 1e-3)
    assert api.sx_range(1, 33, 8, 0, lib=emu_lib) == (1, 5)
    p.close()


def test_stage_timing_survives_more_marks_than_one_flush(emu_lib, tables):
    # with stage timing on, the 8192nd mark flushed, and the closing mark of that flush flushed again
    # (unbounded recursion -> segfault after ~150 HD steps of examples/hd_driver.c).  More than 8192 marks must work,
    # and all of them must be counted.
    p = api.Plan(16, 16, 16, 0, 0, tdir=tables, lib=emu_lib)
    a = p.spectral()
    p.stage_timing(True)
    n = 8192 + 300
    for _ in range(n):
        p.fc_filter(a)
    st = p.stage_times()
    assert sum(c for _, c in st.values()) == n
    p.stage_timing(False)
    p.close()


def test_tuning_knobs_are_validated(emu_lib, tables):
    # a stray value of a tuning variable must fail plan creation instead of selecting an untested variant
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\nfrom specter_b200 import api\n"
            "lib = api.Library(%r)\n"
            "try:\n    api.Plan(16, 16, 64, 25, 5, tdir=%r, lib=lib)\n    print('created')\n"
            "except api.SpecterError as e:\n    print('refused:', e)\n" % (ROOT, emu_lib.path, tables))
    for var, val, ok in (("SX_ZCHUNKS", "0", False), ("SX_XP", "abc", False), ("SX_TMA", "64", False), ("SX_ZCHUNKS", "2", True)):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **{var: val}))
        assert r.returncode == 0, r.stderr
        assert ("created" in r.stdout) == ok and (ok or var in r.stdout), (var, val, r.stdout)
