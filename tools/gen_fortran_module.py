"""Generate the ISO_C_BINDING interface module for include/specter_b200.h.

    python tools/gen_fortran_module.py            # rewrites include/specter_b200_mod.f90
    python tools/gen_fortran_module.py --check    # exit 1 if the committed file is stale

One INTERFACE body per prototype of the header.  The mapping is mechanical:

    sx_plan* / const sx_plan*      TYPE(C_PTR), VALUE               (opaque handle)
    sx_plan**  double**  void**    TYPE(C_PTR), INTENT(OUT)
    const sx_config*               TYPE(sx_config), INTENT(IN)
    int  double  size_t            INTEGER(C_INT) / REAL(C_DOUBLE) / INTEGER(C_SIZE_T), VALUE
    int*  long long*               INTEGER(...), INTENT(OUT)        (scalar results; (*) for the per-stage arrays)
    const char*                    CHARACTER(KIND=C_CHAR), INTENT(IN) :: s(*)   (NUL-terminated: trim(s)//C_NULL_CHAR)
    const double x[n] / double x[n]  REAL(C_DOUBLE), INTENT(IN/OUT) :: x(n)
    double* <scalar result>        REAL(C_DOUBLE), INTENT(OUT)      (out, eng, ens, pot, bot, top, ms of time_end, ...)
    [const] double* *_host, void* host buffers   TYPE(*), DIMENSION(*)   (the driver's own REAL or COMPLEX arrays, Fortran 2018;
                                   OPTIONAL where the header says NULL is skipped: absent = NULL for BIND(C))
    every other [const] double*    TYPE(C_PTR), VALUE               (DEVICE address from sx_malloc / sx_*_state_ptr; C_NULL_PTR = NULL)
    sx_alltoallv_fn / sx_allreduce_fn  TYPE(C_FUNPTR), VALUE        (C_FUNLOC of a BIND(C) wrapper around MPI_Alltoallv / MPI_Allreduce)

No Fortran compiler exists in this image: tests/test_abi.py checks that the module is current, covers every prototype
and keeps the argument counts of the header.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "specter_b200.h")
OUT = os.path.join(ROOT, "include", "specter_b200_mod.f90")

# host-side scalar results passed as double* (everything else that is a double* is a device address)
SCALAR_OUT = {"out", "eng", "ens", "pot", "bot", "top", "bytes_sent"}
# (function, argument) overrides of the name rules
SPECIAL = {
    ("sx_plan_time_end", "ms"): "REAL(C_DOUBLE), INTENT(OUT) :: ms",
    ("sx_plan_comm_stats", "ms"): "REAL(C_DOUBLE), INTENT(OUT) :: ms",
    ("sx_plan_stage_times", "ms"): "REAL(C_DOUBLE), INTENT(OUT) :: ms(*)",
    ("sx_plan_stage_times", "counts"): "INTEGER(C_LONG_LONG), INTENT(OUT) :: counts(*)",
    ("sx_free_host", "hptr"): "TYPE(C_PTR), VALUE :: hptr",
    ("sx_free", "dptr"): "TYPE(C_PTR), VALUE :: dptr",
    ("sx_memcpy_h2d", "dptr"): "TYPE(C_PTR), VALUE :: dptr",
    ("sx_memcpy_d2h", "dptr"): "TYPE(C_PTR), VALUE :: dptr",
    ("sx_plan_set_comm_callbacks", "user"): "TYPE(C_PTR), VALUE :: user",
}
RET = {"int": "INTEGER(C_INT)", "const char*": "TYPE(C_PTR)", "size_t": "INTEGER(C_SIZE_T)",
       "unsigned long long": "INTEGER(C_LONG_LONG)"}


def prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    pat = r"^\s*((?:const\s+)?(?:unsigned\s+long\s+long|int|size_t|char\s*\*))\s+(sx_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;"
    out = []
    for rt, name, args in re.findall(pat, src, flags=re.M | re.S):
        rt = " ".join(rt.split())
        alist = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                m = re.match(r"(.*?)([A-Za-z_0-9]+)(?:\[(\d+)\])?$", a.strip())
                alist.append((m.group(1).strip(), m.group(2), m.group(3)))
        out.append((rt, name, alist))
    return out


def declare(fn, ctype, name, dim):
    if (fn, name) in SPECIAL:
        return SPECIAL[(fn, name)]
    const = ctype.startswith("const ")
    base = ctype[6:] if const else ctype
    if base in ("sx_plan*",):
        return f"TYPE(C_PTR), VALUE :: {name}"
    if base in ("sx_plan**", "double**", "void**"):
        return f"TYPE(C_PTR), INTENT(OUT) :: {name}"
    if base == "sx_config*":
        return f"TYPE(sx_config), INTENT(IN) :: {name}"
    if base == "int":
        return f"INTEGER(C_INT), VALUE :: {name}"
    if base == "size_t":
        return f"INTEGER(C_SIZE_T), VALUE :: {name}"
    if base == "int*":
        return f"INTEGER(C_INT), INTENT(OUT) :: {name}"
    if base == "long long*":
        return f"INTEGER(C_LONG_LONG), INTENT(OUT) :: {name}"
    if base == "char*":
        return f"CHARACTER(KIND=C_CHAR), INTENT(IN) :: {name}(*)"
    if base == "char* const" and dim:
        return f"TYPE(C_PTR), INTENT(IN) :: {name}({dim})"
    if base in ("sx_alltoallv_fn", "sx_allreduce_fn"):
        return f"TYPE(C_FUNPTR), VALUE :: {name}"
    if base == "double" and dim:
        opt = ", OPTIONAL" if name == "b0" else ""     # "b0 may be NULL (no uniform field)"
        return f"REAL(C_DOUBLE), INTENT({'IN' if const else 'OUT'}){opt} :: {name}({dim})"
    if base == "double":
        return f"REAL(C_DOUBLE), VALUE :: {name}"
    if base == "void*":
        return f"TYPE(*), DIMENSION(*){', INTENT(IN)' if const else ''} :: {name}"
    if base == "double*":
        if name.endswith("_host"):
            # "NULL pointers are skipped" (put_state / get_state): an absent OPTIONAL dummy of a BIND(C) procedure is NULL
            opt = ", OPTIONAL" if fn.endswith(("_put_state", "_get_state")) else ""
            return f"TYPE(*), DIMENSION(*){', INTENT(IN)' if const else ''}{opt} :: {name}"
        if name in SCALAR_OUT and not const:
            return f"REAL(C_DOUBLE), INTENT(OUT) :: {name}"
        return f"TYPE(C_PTR), VALUE :: {name}"
    raise SystemExit(f"{fn}: no Fortran mapping for '{ctype} {name}'")


def kinds(decls, ret):
    used = []
    for k in ("C_PTR", "C_FUNPTR", "C_INT", "C_DOUBLE", "C_SIZE_T", "C_LONG_LONG", "C_CHAR", "sx_config"):
        if any(re.search(r"\b%s\b" % k, d) for d in decls + [ret]):
            used.append(k)
    return used


def generate():
    lines = [
        "! specter_b200_mod.f90 -- ISO_C_BINDING interfaces of include/specter_b200.h (libspecter_b200.so).",
        "! GENERATED by tools/gen_fortran_module.py from the header: do not edit, regenerate.",
        "! Goes to src/fftp-b200/ of the reference tree (INTEGRATION.md); link with -lspecter_b200 -lnccl -lcudart.",
        "! Device fields are TYPE(C_PTR) addresses (sx_malloc, sx_*_state_ptr); *_host arguments take the driver's own",
        "! REAL(GP) / COMPLEX(GP) arrays (assumed type, Fortran 2018); strings are NUL-terminated: trim(s)//C_NULL_CHAR.",
        "MODULE specter_b200",
        "  USE, INTRINSIC :: iso_c_binding",
        "  IMPLICIT NONE",
        "  TYPE, BIND(C) :: sx_config",
        "     INTEGER(C_INT) :: nx, ny, nz, Cz, oz, ord",
        "     REAL(C_DOUBLE) :: Lx, Ly, Lz",
        "     TYPE(C_PTR)    :: tdir            ! C_LOC of a NUL-terminated copy of tdir",
        "     INTEGER(C_INT) :: nprocs, myrank, device",
        "  END TYPE sx_config",
        "  INTERFACE",
    ]
    for rt, name, args in prototypes():
        args = [(t, n + "_" if n.lower() in ("end", "function", "type") else n, d) for t, n, d in args]   # keyword-like dummies
        low = [n.lower() for _, n, _ in args]
        if len(set(low)) != len(low):   # Fortran names are case-insensitive
            raise SystemExit(f"{name}: dummy arguments collide case-insensitively: {low}")
        decls = [declare(name, t, n, d) for t, n, d in args]
        ret = RET[rt]
        names = ", ".join(n for _, n, _ in args)
        head = f"     {ret} FUNCTION {name}({names}) BIND(C, NAME='{name}')"
        parts = []
        while len(head) > 120:   # free-form line limit (132): continue after a comma
            cut = head.rfind(", ", 0, 118)
            parts.append(head[:cut + 1] + " &")
            head = "          " + head[cut + 2:]
        lines += parts + [head]
        lines.append("       IMPORT :: " + ", ".join(kinds(decls, ret)))
        lines += ["       " + d for d in decls]
        lines.append(f"     END FUNCTION {name}")
    lines += ["  END INTERFACE", "END MODULE specter_b200", ""]
    return "\n".join(lines)


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    open(OUT, "w").write(text)
    print(OUT, len(prototypes()), "interfaces")
