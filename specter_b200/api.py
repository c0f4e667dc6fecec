"""Host-side mirror of the reference's fftp / pseudo / boundary interface over the C ABI.

The method names and argument meaning follow the Fortran procedures they replace
(``fftp3d_real_to_complex``, ``derivk``, ``gradre``, ``sol_project``,
``v_imposebc_and_project`` ... see include/specter_b200.h for the file:line of each), so
the parity tests read like the reference's own src/tests/*.f90.

There is NO CPU path here: :func:`load_library` only ever opens the nvcc-built
``csrc/libspecter_b200.so`` and raises if it is missing; plan creation fails if no CUDA
device is present.  (tests/ may hand an explicit path of the kernel-emulation build to
:class:`Library` to check kernel logic on CPU -- that build is test infrastructure.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPECTER_B200_LIB") or os.path.join(_HERE, "csrc", "libspecter_b200.so")


class SpecterError(RuntimeError):
    pass


class sx_config(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("Cz", C.c_int), ("oz", C.c_int),
                ("ord", C.c_int), ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("tdir", C.c_char_p), ("nprocs", C.c_int), ("myrank", C.c_int), ("device", C.c_int)]


_P = C.c_void_p
_D = C.c_void_p  # device/host data pointers travel as void*
_I = C.c_int
_F = C.c_double
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)

# name -> argtypes (restype is int unless listed in _RESTYPES); this table is also what
# tests/test_abi.py checks against include/specter_b200.h
SIGNATURES = {
    "sx_last_error": [],
    "sx_version": [],
    "sx_plan_create": [C.POINTER(sx_config), C.POINTER(_P)],
    "sx_plan_destroy": [_P],
    "sx_plan_info": [_P, _PI, _PI, _PI, _PI, _PI],
    "sx_range": [_I, _I, _I, _I, _PI, _PI],
    "sx_plan_launch_count": [_P],
    "sx_plan_synchronize": [_P],
    "sx_plan_release_scratch": [_P],
    "sx_plan_time_begin": [_P],
    "sx_plan_time_end": [_P, _PD],
    "sx_stage_count": [],
    "sx_stage_name": [_I],
    "sx_plan_stage_timing": [_P, _I],
    "sx_plan_stage_times": [_P, _PD, C.POINTER(C.c_longlong), _I],
    "sx_nccl_unique_id": [_D],
    "sx_plan_set_comm": [_P, _D],
    "sx_io_write": [_P, _D, C.c_char_p, C.c_char_p, C.c_char_p],
    "sx_io_read": [_P, _D, C.c_char_p, C.c_char_p, C.c_char_p],
    "sx_hd_output": [_P, C.c_char_p, C.c_char_p, C.c_double, _I],
    "sx_hd_restart": [_P, C.c_char_p, C.c_char_p, C.c_double],
    "sx_output": [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, _I],
    "sx_restart": [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double],
    "sx_benchmark_write": [_P, C.c_char_p, _I, _I, _F, _F, _F],
    "sx_global": [_P, C.c_char_p, C.c_char_p, _I, _F],
    "sx_plan_p2p_export": [_P, _I, _I, _D],
    "sx_plan_p2p_import": [_P, _D],
    "sx_plan_set_comm_callbacks": [_P, _D, _D, _D],
    "sx_plan_comm_stats": [_P, _PD, _PD, C.POINTER(C.c_longlong), _I],
    "sx_malloc": [_P, C.c_size_t, C.POINTER(_D)],
    "sx_free": [_P, _D],
    "sx_malloc_host": [C.c_size_t, C.POINTER(_D)],
    "sx_free_host": [_D],
    "sx_memcpy_h2d": [_P, _D, _D, C.c_size_t],
    "sx_memcpy_d2h": [_P, _D, _D, C.c_size_t],
    "sx_spectral_bytes": [_P],
    "sx_real_bytes": [_P],
    "sx_fftp3d_real_to_complex": [_P, _D, _D],
    "sx_fftp3d_complex_to_real": [_P, _D, _D],
    "sx_fftp2d_real_to_complex_xy": [_P, _D, _D],
    "sx_fftp2d_complex_to_real_xy": [_P, _D, _D],
    "sx_fftp1d_real_to_complex_z": [_P, _D],
    "sx_fftp1d_complex_to_real_z": [_P, _D],
    "sx_derivk": [_P, _D, _D, _I],
    "sx_laplak": [_P, _D, _D],
    "sx_curlk": [_P, _D, _D, _D, _I],
    "sx_fc_filter": [_P, _D],
    "sx_gradre": [_P, _D, _D, _D, _D, _D, _D],
    "sx_prodre": [_P, _D, _D, _D, _D, _D, _D],
    "sx_normvec": [_P, _D, _D, _D, _F, _I],
    "sx_normsca": [_P, _D, _F, _I],
    "sx_normalize": [_P, _D, _D, _D, _F, _I],
    "sx_energy": [_P, _D, _D, _D, _I, _PD],
    "sx_divergence": [_P, _D, _D, _D, _PD],
    "sx_cross": [_P, _D, _D, _D, _D, _D, _D, _I, _PD],
    "sx_hdcheck": [_P, _D, _D, _D, _D, _D, _D, _PD, _PD, _PD],
    "sx_goto_domain_w_boundaries": [_P, _D, _D, _D],
    "sx_goto_3d_fourier": [_P, _D, _D, _D],
    "sx_sol_project": [_P, _D, _D, _D, _D, _I, _I, _I],
    "sx_v_imposebc_and_project": [_P, _D, _D, _D, _D, _I, _PD, _PD],
    "sx_bouncheck_z": [_P, _PD, _PD, _D, _D],
    "sx_vdiagnostic": [_P, _D, _D, _D, _PD],
    "sx_hd_put_state": [_P, _D, _D, _D, _D, _D, _D, _D],
    "sx_hd_get_state": [_P, _D, _D, _D, _D],
    "sx_hd_state_ptr": [_P, _I, C.POINTER(_D)],
    "sx_hd_rkstep1": [_P],
    "sx_hd_rkstep2": [_P, _I, _F, _F, _PD, _PD, _I],
    "sx_hd_step_host": [_P, _D, _D, _D, _D, _D, _D, _D, _F, _F, _PD, _PD],
    "sx_advect": [_P, _D, _D, _D, _D, _D],
    "sx_vector": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _D],
    "sx_variance": [_P, _D, _I, _PD],
    "sx_s_imposebc": [_P, _D],
    "sx_a_imposebc_and_project": [_P, _D, _D, _D, _D],
    "sx_bouss_put_state": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _D],
    "sx_bouss_get_state": [_P, _D, _D, _D, _D, _D],
    "sx_bouss_state_ptr": [_P, _I, C.POINTER(_D)],
    "sx_bouss_rkstep1": [_P],
    "sx_bouss_rkstep2": [_P, _I, _F, _F, _F, _F, _F, _PD, _PD, _I],
    "sx_mhd_put_state": [_P] + [_D] * 13,
    "sx_mhd_get_state": [_P] + [_D] * 8,
    "sx_mhd_state_ptr": [_P, _I, C.POINTER(_D)],
    "sx_mhd_rkstep1": [_P],
    "sx_mhd_rkstep2": [_P, _I, _F, _F, _F, _PD, _I],
    "sx_setup_bc": [_P, C.c_char_p, C.POINTER(C.c_char_p)],
    "sx_neumann_reconstruct": [_P, _D, _I, _I],
    "sx_robin_reconstruct": [_P, _D, _I, _D],
    "sx_helicity": [_P, _D, _D, _D, _PD],
    "sx_product": [_P, _D, _D, _PD],
    "sx_pscheck": [_P, _D, _D, _PD],
    "sx_maxabs": [_P, _D, _D, _D, _I, _PD],
    "sx_mhdcheck": [_P, _D, _D, _D, _D, _D, _D, _I, _I, _PD],
    "sx_robcheck": [_P, _D, _D, _D, _PD],
    "sx_bdiagnostic": [_P, _D, _D, _D, _PD, _PD, _PI],
    "sx_sdiagnostic": [_P, _D, _PD],
    "sx_rotbouss_rkstep2": [_P, _I, _F, _F, _F, _F, _F, _PD, _PD, _PD, _I],
    "sx_mhdbouss_put_state": [_P] + [_D] * 15,
    "sx_mhdbouss_get_state": [_P] + [_D] * 9,
    "sx_mhdbouss_state_ptr": [_P, _I, C.POINTER(_D)],
    "sx_mhdbouss_rkstep1": [_P],
    "sx_mhdbouss_rkstep2": [_P, _I, _F, _F, _F, _F, _F, _F, _PD, _I],
    "sx_boots_points": [_I, _I, _PI, _PI],
    "sx_boots_regrid": [_I, _I, _I, _I, _I, C.c_char_p, _I, _I, _I, _D, _D],
    "sx_boots_files": [_I, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _I, _I, _I, _I, _I, _I, _I],
    "sx_boots_launch_count": [],
}
_RESTYPES = {"sx_stage_name": C.c_char_p, "sx_last_error": C.c_char_p, "sx_version": C.c_char_p,
             "sx_plan_launch_count": C.c_ulonglong, "sx_boots_launch_count": C.c_ulonglong, "sx_spectral_bytes": C.c_size_t,
             "sx_real_bytes": C.c_size_t}


def _preload_nccl():
    """The library links libnccl.so.2; PyTorch ships its own (newer) copy.  Whichever is mapped first serves both,
    and torch's CUDA library needs symbols of its own version: map that copy first when it is installed, so that
    the import order of torch and specter_b200 does not matter."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        cand = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


class Library:
    """ctypes binding of one build of the C ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise SpecterError(
                f"{path} not found: build the CUDA extension first (python -m specter_b200.build); "
                "specter_b200 has no CPU fallback")
        self.path = path
        _preload_nccl()
        self.dll = C.CDLL(path)
        for name, args in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the ABI is incomplete
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)

    def check(self, rc: int):
        if rc != 0:
            raise SpecterError(self.dll.sx_last_error().decode())

    def version(self) -> str:
        return self.dll.sx_version().decode()


_LIB: Optional[Library] = None


def load_library() -> Library:
    """The product library (CUDA, sm_100a).  Never falls back to anything else."""
    global _LIB
    if _LIB is None:
        lib = Library(LIB_PATH)
        if "sm_100a" not in lib.version():   # e.g. SPECTER_B200_LIB pointing at the tests' kernel-emulation build
            raise SpecterError(f"{LIB_PATH} is not the nvcc-built sm_100a library ({lib.version()}); "
                               "specter_b200 has no CPU fallback")
        _LIB = lib
    return _LIB


def sx_range(n1, n2, nprocs, irank, lib: Optional[Library] = None):
    """``range`` (fftp.fpp:1154-1184)."""
    lib = lib or load_library()
    a, b = C.c_int(), C.c_int()
    lib.check(lib.dll.sx_range(n1, n2, nprocs, irank, C.byref(a), C.byref(b)))
    return a.value, b.value


# ---- BOOTS regridder (tools/boots.fpp) ----
def boots_points(nzt, nzp, lib: Optional[Library] = None):
    """Continuation points (Czt, Czn) of the old / new grid (boots.fpp:176-182); nzp = nz-Cz of the new grid."""
    lib = lib or load_library()
    a, b = C.c_int(), C.c_int()
    lib.check(lib.dll.sx_boots_points(int(nzt), int(nzp), C.byref(a), C.byref(b)))
    return a.value, b.value


def boots_regrid(vt: np.ndarray, nx, ny, nzp, ozt, tdir, device=-1, lib: Optional[Library] = None) -> np.ndarray:
    """One field of the BOOTS3D loop (boots.fpp:249-303): vt[k,j,i] on the old physical grid (nzt, nyt, nxt)
    -> the field on (nzp, ny, nx)."""
    lib = lib or load_library()
    vt = np.ascontiguousarray(vt, dtype=np.float64)
    nzt, nyt, nxt = vt.shape
    out = np.empty((int(nzp), int(ny), int(nx)), dtype=np.float64)
    lib.check(lib.dll.sx_boots_regrid(int(device), nxt, nyt, nzt, int(ozt), str(tdir).encode(), int(nx), int(ny), int(nzp),
                                      vt.ctypes.data, out.ctypes.data))
    return out


def boots_files(idir, odir, tdir, fnlist, nxt, nyt, nzt, ozt, nx, ny, nzp, device=-1, lib: Optional[Library] = None):
    """The `regrid' / `order' namelists of boots.inp as arguments: every file of fnlist (';' separated) in idir is
    prolongated and written to odir with the suffix _P<nx>-<ny>-<nzp> (boots.fpp:174, 228-312)."""
    lib = lib or load_library()
    lib.check(lib.dll.sx_boots_files(int(device), str(idir).encode(), str(odir).encode(), str(tdir).encode(),
                                     str(fnlist).encode(), int(nxt), int(nyt), int(nzt), int(ozt), int(nx), int(ny), int(nzp)))


def boots_launch_count(lib: Optional[Library] = None) -> int:
    lib = lib or load_library()
    return int(lib.dll.sx_boots_launch_count())


class DeviceArray:
    """A device buffer in one of the two reference layouts."""

    def __init__(self, plan: "Plan", kind: str):
        self.plan = plan
        self.kind = kind  # "spectral" (complex128 (nxl,ny,nz) C-order == Fortran (nz,ny,ista:iend)) | "real"
        self.shape = plan.cshape if kind == "spectral" else plan.rshape
        self.dtype = np.complex128 if kind == "spectral" else np.float64
        self.nbytes = int(np.prod(self.shape)) * np.dtype(self.dtype).itemsize
        ptr = C.c_void_p()
        plan.lib.check(plan.lib.dll.sx_malloc(plan.handle, max(self.nbytes, 16), C.byref(ptr)))
        self.ptr = ptr
        self.owned = True

    @classmethod
    def view(cls, plan: "Plan", kind: str, ptr) -> "DeviceArray":
        """Wrap plan-owned device memory (not freed by this object)."""
        self = cls.__new__(cls)
        self.plan, self.kind = plan, kind
        self.shape = plan.cshape if kind == "spectral" else plan.rshape
        self.dtype = np.complex128 if kind == "spectral" else np.float64
        self.nbytes = int(np.prod(self.shape)) * np.dtype(self.dtype).itemsize
        self.ptr = ptr
        self.owned = False
        return self

    def put(self, host: np.ndarray) -> "DeviceArray":
        h = np.ascontiguousarray(host, dtype=self.dtype)
        if h.shape != tuple(self.shape):
            raise SpecterError(f"shape mismatch: got {h.shape}, expected {tuple(self.shape)}")
        self.plan.lib.check(self.plan.lib.dll.sx_memcpy_h2d(self.plan.handle, self.ptr, h.ctypes.data, self.nbytes))
        return self

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        self.plan.lib.check(self.plan.lib.dll.sx_memcpy_d2h(self.plan.handle, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def free(self):
        if self.ptr and self.owned and self.plan.handle:
            self.plan.lib.dll.sx_free(self.plan.handle, self.ptr)
        self.ptr = None


def _vec2(v):
    return (C.c_double * 2)(float(v[0]), float(v[1]))


class Plan:
    """FCPLAN + BCPLAN + grid of the reference, bound to one GPU (fcgram_create_plan, setup_bc)."""

    def __init__(self, nx, ny, nz, Cz, oz, ord=2, Lx=1.0, Ly=1.0, Lz=1.0, tdir="", nprocs=1, myrank=0,
                 device=-1, lib: Optional[Library] = None):
        self.lib = lib or load_library()
        self.p2p = False
        self.cfg = sx_config(nx, ny, nz, Cz, oz, ord, Lx, Ly, Lz, tdir.encode(), nprocs, myrank, device)
        self.handle = C.c_void_p()
        self.lib.check(self.lib.dll.sx_plan_create(C.byref(self.cfg), C.byref(self.handle)))
        v = [C.c_int() for _ in range(5)]
        self.lib.check(self.lib.dll.sx_plan_info(self.handle, *[C.byref(x) for x in v]))
        self.ista, self.iend, self.ksta, self.kend, self.pkend = [x.value for x in v]
        self.nx, self.ny, self.nz, self.Cz, self.oz, self.ord = nx, ny, nz, Cz, oz, ord
        self.nxl = self.iend - self.ista + 1
        self.nzl = self.kend - self.ksta + 1
        self.cshape = (self.nxl, ny, nz)
        self.rshape = (self.nzl, ny, nx)
        self._arrays = []
        self._pinned = []

    # ---- memory ----
    def spectral(self, host: Optional[np.ndarray] = None) -> DeviceArray:
        a = DeviceArray(self, "spectral")
        self._arrays.append(a)
        return a.put(host) if host is not None else a

    def real(self, host: Optional[np.ndarray] = None) -> DeviceArray:
        a = DeviceArray(self, "real")
        self._arrays.append(a)
        return a.put(host) if host is not None else a

    def pinned_like(self, a: np.ndarray) -> np.ndarray:
        """A page-locked host copy of `a` (cudaMallocHost through the C ABI), freed by close()."""
        ptr = C.c_void_p()
        self.lib.check(self.lib.dll.sx_malloc_host(max(a.nbytes, 16), C.byref(ptr)))
        self._pinned.append(ptr)
        buf = (C.c_char * a.nbytes).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
        out[...] = a
        return out

    def pinned_empty(self) -> np.ndarray:
        """An uninitialised page-locked host array of the plan's spectral shape (freed by close())."""
        n = int(np.prod(self.cshape)) * 16
        ptr = C.c_void_p()
        self.lib.check(self.lib.dll.sx_malloc_host(max(n, 16), C.byref(ptr)))
        self._pinned.append(ptr)
        buf = (C.c_char * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.complex128).reshape(self.cshape)

    # ---- multi-GPU ----
    def init_comm_torch(self, dist, p2p=True, p2p_fields=(14, 7)):
        """Create the plan's own NCCL communicator; the 128-byte unique id travels over the caller's
        torch.distributed group (the Fortran driver would MPI_BCAST it).  p2p_fields = (inverse, forward) transposed
        fields the peer-to-peer arena is sized for: HD 6/3, BOUSS 8/4, MHD 12/6, MHDBOUSS 14/7 (the default)."""
        import torch
        buf = (C.c_char * 128)()
        if dist.get_rank() == 0:
            self.lib.check(self.lib.dll.sx_nccl_unique_id(buf))
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        self._call("sx_plan_set_comm", C.c_char_p(raw))
        if p2p and dist.get_backend() == "nccl" and os.environ.get("SX_P2P", "1") != "0":
            self.init_p2p_torch(dist, *p2p_fields)

    def init_p2p_torch(self, dist, n_inverse=6, n_forward=3):
        """Peer-to-peer exchange buffers: export this rank's receive arena, all-gather the 64-byte CUDA IPC
        handles over the caller's group, map the peers' arenas (sx_plan_p2p_export / sx_plan_p2p_import)."""
        import torch
        buf = (C.c_char * 64)()
        self._call("sx_plan_p2p_export", n_inverse, n_forward, buf)
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        allh = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allh, mine)
        raw = b"".join(bytes(h.cpu().tolist()) for h in allh)
        self._call("sx_plan_p2p_import", C.c_char_p(raw))
        self.p2p = True

    def set_comm_callbacks(self, alltoallv, allreduce):
        """Route the slab exchange through caller code: alltoallv(send_ptr, sdispl, scount, recv_ptr,
        rdispl, rcount) with byte offsets per rank, allreduce(ptr, n) summing n host doubles in place."""
        A2A = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p,
                          C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int)
        ARD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int)

        def a2a(user, send, sd, sc, recv, rd, rc, n):
            try:
                alltoallv(send, [sd[i] for i in range(n)], [sc[i] for i in range(n)], recv,
                          [rd[i] for i in range(n)], [rc[i] for i in range(n)])
                return 0
            except Exception as e:  # pragma: no cover
                print("alltoallv callback failed:", e)
                return 1

        def ard(user, ptr, n):
            try:
                allreduce(ptr, n)
                return 0
            except Exception as e:  # pragma: no cover
                print("allreduce callback failed:", e)
                return 1

        self._cb = (A2A(a2a), ARD(ard))  # keep alive
        self._call("sx_plan_set_comm_callbacks", C.cast(self._cb[0], C.c_void_p), C.cast(self._cb[1], C.c_void_p), None)

    def comm_stats(self, reset=False):
        b, ms, n = C.c_double(), C.c_double(), C.c_longlong()
        self._call("sx_plan_comm_stats", C.byref(b), C.byref(ms), C.byref(n), 1 if reset else 0)
        return {"bytes_sent": b.value, "ms": ms.value, "exchanges": n.value}

    def close(self):
        if self.handle:
            for q in self._pinned:
                self.lib.dll.sx_free_host(q)
            self._pinned = []
            for a in self._arrays:
                a.free()
            self._arrays = []
            self.lib.dll.sx_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def release_scratch(self):
        """Free the pooled temporaries of the per-operator entries (re-grown on demand)."""
        self._call("sx_plan_release_scratch")

    def synchronize(self):
        self.lib.check(self.lib.dll.sx_plan_synchronize(self.handle))

    def time_begin(self):
        self._call("sx_plan_time_begin")

    def time_end(self) -> float:
        """Elapsed device milliseconds on the plan's stream since time_begin (synchronises)."""
        ms = C.c_double()
        self._call("sx_plan_time_end", C.byref(ms))
        return ms.value

    def stage_timing(self, on: bool):
        self._call("sx_plan_stage_timing", 1 if on else 0)

    def stage_times(self) -> dict:
        """{stage name: (milliseconds, launches)} accumulated since stage_timing(True)."""
        n = self.lib.dll.sx_stage_count()
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self._call("sx_plan_stage_times", ms, cnt, n)
        return {self.lib.dll.sx_stage_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}

    @property
    def launch_count(self) -> int:
        return int(self.lib.dll.sx_plan_launch_count(self.handle))

    def _call(self, name, *args):
        self.lib.check(getattr(self.lib.dll, name)(self.handle, *args))

    # ---- fftp ----
    def fftp3d_real_to_complex(self, r: DeviceArray, out: DeviceArray):
        self._call("sx_fftp3d_real_to_complex", r.ptr, out.ptr)

    def fftp3d_complex_to_real(self, a: DeviceArray, out: DeviceArray):
        self._call("sx_fftp3d_complex_to_real", a.ptr, out.ptr)

    def fftp2d_real_to_complex_xy(self, r: DeviceArray, out: DeviceArray):
        self._call("sx_fftp2d_real_to_complex_xy", r.ptr, out.ptr)

    def fftp2d_complex_to_real_xy(self, a: DeviceArray, out: DeviceArray):
        self._call("sx_fftp2d_complex_to_real_xy", a.ptr, out.ptr)

    def fftp1d_real_to_complex_z(self, a: DeviceArray):
        self._call("sx_fftp1d_real_to_complex_z", a.ptr)

    def fftp1d_complex_to_real_z(self, a: DeviceArray):
        self._call("sx_fftp1d_complex_to_real_z", a.ptr)

    # ---- pseudo ----
    def derivk(self, a, b, dir):
        self._call("sx_derivk", a.ptr, b.ptr, dir)

    def laplak(self, a, b):
        self._call("sx_laplak", a.ptr, b.ptr)

    def curlk(self, a, b, c, dir):
        self._call("sx_curlk", a.ptr, b.ptr, c.ptr, dir)

    def fc_filter(self, a):
        self._call("sx_fc_filter", a.ptr)

    def gradre(self, a, b, c, d, e, f):
        self._call("sx_gradre", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, f.ptr)

    def prodre(self, a, b, c, d, e, f):
        self._call("sx_prodre", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, f.ptr)

    def normvec(self, a, b, c, d, kin):
        self._call("sx_normvec", a.ptr, b.ptr, c.ptr, float(d), int(kin))

    def normsca(self, a, b, kin):
        self._call("sx_normsca", a.ptr, float(b), int(kin))

    def normalize(self, fx, fy, fz, f0, kin):
        self._call("sx_normalize", fx.ptr, fy.ptr, fz.ptr, float(f0), int(kin))

    def energy(self, a, b, c, kin) -> float:
        out = C.c_double()
        self._call("sx_energy", a.ptr, b.ptr, c.ptr, kin, C.byref(out))
        return out.value

    def divergence(self, a, b, c) -> float:
        out = C.c_double()
        self._call("sx_divergence", a.ptr, b.ptr, c.ptr, C.byref(out))
        return out.value

    def cross(self, a, b, c, d, e, f, kin) -> float:
        out = C.c_double()
        self._call("sx_cross", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, f.ptr, kin, C.byref(out))
        return out.value

    def hdcheck(self, a, b, c, d, e, f):
        o = [C.c_double() for _ in range(3)]
        self._call("sx_hdcheck", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, f.ptr, *[C.byref(x) for x in o])
        return tuple(x.value for x in o)

    # ---- boundary ----
    def goto_domain_w_boundaries(self, a, b=None, c=None):
        self._call("sx_goto_domain_w_boundaries", a.ptr, b.ptr if b is not None else None, c.ptr if c is not None else None)

    def goto_3d_fourier(self, a, b=None, c=None):
        self._call("sx_goto_3d_fourier", a.ptr, b.ptr if b is not None else None, c.ptr if c is not None else None)

    def sol_project(self, a, b, c, d, bctarget, bczsta, bczend):
        self._call("sx_sol_project", a.ptr, b.ptr, c.ptr, d.ptr, bctarget, bczsta, bczend)

    def v_imposebc_and_project(self, vx, vy, vz, pr, rki, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
        self._call("sx_v_imposebc_and_project", vx.ptr, vy.ptr, vz.ptr, pr.ptr, rki, _vec2(v_zsta), _vec2(v_zend))

    def bouncheck_z(self, a, b=None):
        bot, top = C.c_double(), C.c_double()
        self._call("sx_bouncheck_z", C.byref(bot), C.byref(top), a.ptr, b.ptr if b is not None else None)
        return bot.value, top.value

    def vdiagnostic(self, a, b, c):
        out = (C.c_double * 5)()
        self._call("sx_vdiagnostic", a.ptr, b.ptr, c.ptr, out)
        return tuple(out)

    # ---- HD substep on plan-owned state ----
    def hd_put_state(self, vx=None, vy=None, vz=None, pr=None, fx=None, fy=None, fz=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.complex128)
                for a in (vx, vy, vz, pr, fx, fy, fz)]
        for a in arrs:
            if a is not None and a.shape != tuple(self.cshape):
                raise SpecterError("hd_put_state: shape mismatch")
        self._call("sx_hd_put_state", *[None if a is None else a.ctypes.data for a in arrs])

    def hd_get_state(self):
        out = [np.empty(self.cshape, dtype=np.complex128) for _ in range(4)]
        self._call("sx_hd_get_state", *[a.ctypes.data for a in out])
        return out

    def hd_field(self, which: int) -> DeviceArray:
        """Plan-owned HD state as a device array: 0..2 v, 3 pr, 4..6 f, 7..9 RK base."""
        ptr = C.c_void_p()
        self._call("sx_hd_state_ptr", which, C.byref(ptr))
        return DeviceArray.view(self, "spectral", ptr)

    def hd_rkstep1(self):
        self._call("sx_hd_rkstep1")

    def hd_rkstep2(self, o, dt, nu, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0), impl=0):
        self._call("sx_hd_rkstep2", o, dt, nu, _vec2(v_zsta), _vec2(v_zend), impl)

    def hd_step(self, dt, nu, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0), impl=0):
        """rkstep1 + ord substeps on the device-resident state (specter.fpp:1142-1161)."""
        self.hd_rkstep1()
        for o in range(self.ord, 0, -1):
            self.hd_rkstep2(o, dt, nu, v_zsta, v_zend, impl)

    def io_write(self, r: DeviceArray, dir, fname, nmb):
        self._call("sx_io_write", r.ptr, str(dir).encode(), fname.encode(), nmb.encode())

    def io_read(self, r: DeviceArray, dir, fname, nmb):
        self._call("sx_io_read", r.ptr, str(dir).encode(), fname.encode(), nmb.encode())

    def hd_output(self, odir, ext, dt, outs=0):
        self._call("sx_hd_output", str(odir).encode(), ext.encode(), float(dt), int(outs))

    def hd_restart(self, idir, ext, dt):
        self._call("sx_hd_restart", str(idir).encode(), ext.encode(), float(dt))

    def hd_step_host(self, vx, vy, vz, pr, fx, fy, fz, dt, nu, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0)):
        """One full step on HOST arrays (in place): H2D, rkstep1 + ord substeps, D2H.  fx / fy / fz may be None:
        the forcing of an earlier call (or of hd_put_state) is kept on the device."""
        for a in (vx, vy, vz, pr, fx, fy, fz):
            if a is not None and not (a.flags["C_CONTIGUOUS"] and a.dtype == np.complex128 and a.shape == tuple(self.cshape)):
                raise SpecterError("hd_step_host needs C-contiguous complex128 arrays of the plan's spectral shape")
        if any(a is None for a in (vx, vy, vz, pr)):
            raise SpecterError("hd_step_host: vx, vy, vz, pr are required")
        self._call("sx_hd_step_host", vx.ctypes.data, vy.ctypes.data, vz.ctypes.data, pr.ctypes.data,
                   *[None if a is None else a.ctypes.data for a in (fx, fy, fz)], dt, nu, _vec2(v_zsta), _vec2(v_zend))

    def advect(self, a, b, c, d, e):
        self._call("sx_advect", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr)

    def vector(self, a, b, c, d, e, f, x, y, z):
        self._call("sx_vector", a.ptr, b.ptr, c.ptr, d.ptr, e.ptr, f.ptr, x.ptr, y.ptr, z.ptr)

    def variance(self, a, kin) -> float:
        out = C.c_double()
        self._call("sx_variance", a.ptr, kin, C.byref(out))
        return out.value

    def s_imposebc(self, th):
        self._call("sx_s_imposebc", th.ptr)

    def a_imposebc_and_project(self, ax, ay, az, ph):
        self._call("sx_a_imposebc_and_project", ax.ptr, ay.ptr, az.ptr, ph.ptr)

    def _host_fields(self, arrs, what):
        out = [None if a is None else np.ascontiguousarray(a, dtype=np.complex128) for a in arrs]
        for a in out:
            if a is not None and a.shape != tuple(self.cshape):
                raise SpecterError(f"{what}: shape mismatch")
        return out

    # ---- Boussinesq substep on plan-owned state (include/bouss/bouss_rkstep{1,2}.f90) ----
    def bouss_put_state(self, vx=None, vy=None, vz=None, pr=None, th=None, fx=None, fy=None, fz=None, fs=None):
        arrs = self._host_fields((vx, vy, vz, pr, th, fx, fy, fz, fs), "bouss_put_state")
        self._call("sx_bouss_put_state", *[None if a is None else a.ctypes.data for a in arrs])

    def bouss_get_state(self):
        out = [np.empty(self.cshape, dtype=np.complex128) for _ in range(5)]
        self._call("sx_bouss_get_state", *[a.ctypes.data for a in out])
        return out

    def bouss_field(self, which: int) -> DeviceArray:
        ptr = C.c_void_p()
        self._call("sx_bouss_state_ptr", which, C.byref(ptr))
        return DeviceArray.view(self, "spectral", ptr)

    def bouss_rkstep1(self):
        self._call("sx_bouss_rkstep1")

    def bouss_rkstep2(self, o, dt, nu, kappa, xmom=1.0, xtemp=1.0, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0), impl=0):
        self._call("sx_bouss_rkstep2", o, dt, nu, kappa, xmom, xtemp, _vec2(v_zsta), _vec2(v_zend), impl)

    def bouss_step(self, dt, nu, kappa, xmom=1.0, xtemp=1.0, v_zsta=(0.0, 0.0), v_zend=(0.0, 0.0), impl=0):
        self.bouss_rkstep1()
        for o in range(self.ord, 0, -1):
            self.bouss_rkstep2(o, dt, nu, kappa, xmom, xtemp, v_zsta, v_zend, impl)

    # ---- MHD substep on plan-owned state (include/mhd/mhd_rkstep{1,2}.f90) ----
    def mhd_put_state(self, vx=None, vy=None, vz=None, pr=None, ax=None, ay=None, az=None, fx=None, fy=None, fz=None,
                      mx=None, my=None, mz=None):
        arrs = self._host_fields((vx, vy, vz, pr, ax, ay, az, fx, fy, fz, mx, my, mz), "mhd_put_state")
        self._call("sx_mhd_put_state", *[None if a is None else a.ctypes.data for a in arrs])

    def mhd_get_state(self):
        out = [np.empty(self.cshape, dtype=np.complex128) for _ in range(8)]
        self._call("sx_mhd_get_state", *[a.ctypes.data for a in out])
        return out

    def mhd_field(self, which: int) -> DeviceArray:
        ptr = C.c_void_p()
        self._call("sx_mhd_state_ptr", which, C.byref(ptr))
        return DeviceArray.view(self, "spectral", ptr)

    def mhd_rkstep1(self):
        self._call("sx_mhd_rkstep1")

    def mhd_rkstep2(self, o, dt, nu, mu, b0=(0.0, 0.0, 0.0), impl=0):
        b = (C.c_double * 3)(*[float(x) for x in b0])
        self._call("sx_mhd_rkstep2", o, dt, nu, mu, b, impl)

    def mhd_step(self, dt, nu, mu, b0=(0.0, 0.0, 0.0), impl=0):
        self.mhd_rkstep1()
        for o in range(self.ord, 0, -1):
            self.mhd_rkstep2(o, dt, nu, mu, b0, impl)

    # ---- wall BC kinds, stand-alone reconstructions, remaining diagnostics ----
    def setup_bc(self, field, bckind):
        """``setup_bc`` (boundary_mod.fpp:30-68): field 'v' | 's' | 'b', bckind = six strings as in parameter.inp."""
        if len(bckind) != 6:
            raise SpecterError("setup_bc: bckind takes six strings (x0, xL, y0, yL, z0, zL)")
        arr = (C.c_char_p * 6)(*[str(b).encode() for b in bckind])
        self._call("sx_setup_bc", field.encode(), arr)

    def neumann_reconstruct(self, f, boun, order):
        self._call("sx_neumann_reconstruct", f.ptr, boun, order)

    def robin_reconstruct(self, f, boun, a=None):
        """a: None for khom (what the reference's callers pass), or a device pointer of (nxl, ny) real coefficients."""
        self._call("sx_robin_reconstruct", f.ptr, boun, a)

    def _scalar_out(self, name, *args):
        out = C.c_double()
        self._call(name, *args, C.byref(out))
        return out.value

    def helicity(self, a, b, c) -> float:
        return self._scalar_out("sx_helicity", a.ptr, b.ptr, c.ptr)

    def product(self, a, b) -> float:
        return self._scalar_out("sx_product", a.ptr, b.ptr)

    def pscheck(self, a, b):
        out = (C.c_double * 3)()
        self._call("sx_pscheck", a.ptr, b.ptr, out)
        return tuple(out)

    def maxabs(self, a, b, c, kin) -> float:
        return self._scalar_out("sx_maxabs", a.ptr, b.ptr, c.ptr, kin)

    def mhdcheck(self, a, b, c, ma, mb, mc, hel=1, crs=1):
        out = (C.c_double * 9)()
        self._call("sx_mhdcheck", a.ptr, b.ptr, c.ptr, ma.ptr, mb.ptr, mc.ptr, hel, crs, out)
        return tuple(out)

    def robcheck(self, a, b, c):
        out = (C.c_double * 4)()
        self._call("sx_robcheck", a.ptr, b.ptr, c.ptr, out)
        return tuple(out)

    def bdiagnostic(self, a, b, c) -> dict:
        cond, vac, which = (C.c_double * 6)(), (C.c_double * 6)(), C.c_int()
        self._call("sx_bdiagnostic", a.ptr, b.ptr, c.ptr, cond, vac, C.byref(which))
        out = {}
        if which.value & 1:
            out["conducting"] = tuple(cond)
        if which.value & 2:
            out["vacuum"] = tuple(vac)
        return out

    def sdiagnostic(self, a):
        out = (C.c_double * 2)()
        self._call("sx_sdiagnostic", a.ptr, out)
        return tuple(out)

    # ---- ROTBOUSS on the BOUSS state (include/rotbouss/rotbouss_rkstep{1,2}.f90) ----
    def rotbouss_rkstep2(self, o, dt, nu, kappa, xmom=1.0, xtemp=1.0, omega=(0.0, 0.0, 0.0), v_zsta=(0.0, 0.0),
                         v_zend=(0.0, 0.0), impl=0):
        om = (C.c_double * 3)(*[float(x) for x in omega])
        self._call("sx_rotbouss_rkstep2", o, dt, nu, kappa, xmom, xtemp, om, _vec2(v_zsta), _vec2(v_zend), impl)

    def rotbouss_step(self, dt, nu, kappa, xmom=1.0, xtemp=1.0, omega=(0.0, 0.0, 0.0), v_zsta=(0.0, 0.0),
                      v_zend=(0.0, 0.0), impl=0):
        self.bouss_rkstep1()
        for o in range(self.ord, 0, -1):
            self.rotbouss_rkstep2(o, dt, nu, kappa, xmom, xtemp, omega, v_zsta, v_zend, impl)

    # ---- MHDBOUSS on plan-owned state (include/mhdbouss/mhdbouss_rkstep{1,2}.f90) ----
    def mhdbouss_put_state(self, vx=None, vy=None, vz=None, pr=None, ax=None, ay=None, az=None, th=None, fx=None,
                           fy=None, fz=None, mx=None, my=None, mz=None, fs=None):
        arrs = self._host_fields((vx, vy, vz, pr, ax, ay, az, th, fx, fy, fz, mx, my, mz, fs), "mhdbouss_put_state")
        self._call("sx_mhdbouss_put_state", *[None if a is None else a.ctypes.data for a in arrs])

    def mhdbouss_get_state(self):
        out = [np.empty(self.cshape, dtype=np.complex128) for _ in range(9)]
        self._call("sx_mhdbouss_get_state", *[a.ctypes.data for a in out])
        return out

    def mhdbouss_field(self, which: int) -> DeviceArray:
        ptr = C.c_void_p()
        self._call("sx_mhdbouss_state_ptr", which, C.byref(ptr))
        return DeviceArray.view(self, "spectral", ptr)

    def mhdbouss_rkstep1(self):
        self._call("sx_mhdbouss_rkstep1")

    def mhdbouss_rkstep2(self, o, dt, nu, mu, kappa, xmom=1.0, xtemp=1.0, b0=(0.0, 0.0, 0.0), impl=0):
        b = (C.c_double * 3)(*[float(x) for x in b0])
        self._call("sx_mhdbouss_rkstep2", o, dt, nu, mu, kappa, xmom, xtemp, b, impl)

    def mhdbouss_step(self, dt, nu, mu, kappa, xmom=1.0, xtemp=1.0, b0=(0.0, 0.0, 0.0), impl=0):
        self.mhdbouss_rkstep1()
        for o in range(self.ord, 0, -1):
            self.mhdbouss_rkstep2(o, dt, nu, mu, kappa, xmom, xtemp, b0, impl)

    # ---- output / restart blocks for any solver, benchmark.txt ----
    def output(self, solver, odir, ext, dt, outs=0):
        self._call("sx_output", solver.encode(), str(odir).encode(), ext.encode(), float(dt), int(outs))

    def restart(self, solver, idir, ext, dt):
        self._call("sx_restart", solver.encode(), str(idir).encode(), ext.encode(), float(dt))

    def global_quantities(self, solver, odir, t, dt):
        """The `<solver>_global.f90` include: appends the rows of balance.txt & co. in the reference's formats."""
        self._call("sx_global", solver.encode(), str(odir).encode(), int(t), float(dt))

    def benchmark_write(self, path, nsteps, nth=1, tcpu=0.0, tomp=0.0, twtime=0.0):
        self._call("sx_benchmark_write", str(path).encode(), int(nsteps), int(nth), float(tcpu), float(tomp), float(twtime))
