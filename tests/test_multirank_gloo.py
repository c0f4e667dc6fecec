"""world_size-2 / 3 CPU test of the slab-parallel fused substep: every rank runs the kernel sources
under the CPU emulation on its own kx-slab; the all-to-all-v blocks travel through
torch.distributed (gloo) via sx_plan_set_comm_callbacks -- the same block tables NCCL uses on GPUs.
The gathered result must equal the single-rank oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, shape, ord_, emu_path, tables, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import specter_oracle as O
    from specter_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = api.Library(emu_path)
        nx, ny, nz = shape
        p = api.Plan(nx, ny, nz, 25, 5, ord=ord_, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, nprocs=world, myrank=rank, lib=lib)

        def alltoallv(send, sd, sc, recv, rd, rc):
            assert os.environ.get("SX_TEST_P2P") != "1", "the fused substep must not use the all-to-all callback with the peer-to-peer transport"
            ins = [torch.from_numpy(np.frombuffer((C.c_char * sc[r]).from_address(send + sd[r]), dtype=np.uint8).copy())
                   if sc[r] else torch.empty(0, dtype=torch.uint8) for r in range(world)]
            outs = [torch.empty(rc[r], dtype=torch.uint8) for r in range(world)]
            dist.all_to_all(outs, ins) if dist.get_backend() != "gloo" else _gloo_a2a(dist, outs, ins, rank, world)
            for r in range(world):
                if rc[r]:
                    C.memmove(recv + rd[r], outs[r].numpy().ctypes.data, rc[r])

        def allreduce(ptr, n):
            t = torch.tensor([ptr[i] for i in range(n)], dtype=torch.float64)
            dist.all_reduce(t)
            for i in range(n):
                ptr[i] = float(t[i])

        p.set_comm_callbacks(alltoallv, allreduce)
        _maybe_p2p(p, dist, 6, 3)
        # single-rank oracle state, sliced to this rank's kx-slab
        g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=ord_)
        s = O.make_hd_state(g)
        sl = slice(p.ista - 1, p.iend)
        p.hd_put_state(*[np.ascontiguousarray(a[sl]) for a in (s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)])
        p.hd_step(1e-3, 1e-3)
        got = p.hd_get_state()
        v = [p.hd_field(i) for i in range(3)]
        eng = p.energy(*v, 1)
        div = p.divergence(*v)
        O.hd_step(g, s, 1e-3, 1e-3)
        scale = max(np.abs(a).max() for a in (s.vx, s.vy, s.vz))
        err = max(np.abs(a - b[sl]).max() for a, b in zip(got[:3], (s.vx, s.vy, s.vz))) / scale
        q.put((rank, float(err), eng, O.energy(g, s.vx, s.vy, s.vz, 1), div, O.divergence(g, s.vx, s.vy, s.vz),
               (p.ista, p.iend)))
        p.close()
    finally:
        dist.destroy_process_group()


def _maybe_p2p(p, dist, n_inverse, n_forward):
    """SX_TEST_P2P=1: the DEFAULT transport of the GPU build on one node -- peer-to-peer copies into the peers' receive
    arenas (here POSIX shared memory between the rank processes instead of CUDA IPC) with the chunked xy pipeline, the
    callbacks only carrying the completion barrier."""
    if os.environ.get("SX_TEST_P2P") == "1":
        p.init_p2p_torch(dist, n_inverse, n_forward)
        assert p.p2p


def _gloo_a2a(dist, outs, ins, rank, world):
    """gloo has no all_to_all: pairwise exchange."""
    reqs = []
    for r in range(world):
        if r == rank:
            outs[r].copy_(ins[r])
            continue
        if ins[r].numel():
            reqs.append(dist.isend(ins[r], r))
        if outs[r].numel():
            reqs.append(dist.irecv(outs[r], r))
    for q in reqs:
        q.wait()


@pytest.mark.parametrize("world,shape,ord_,tma_min", [(2, (32, 16, 64), 2, None), (3, (16, 16, 64), 2, None),
                                                      (2, (16, 128, 128), 2, "16")])
def test_fused_substep_multirank(world, shape, ord_, tma_min, emu_lib, tables, monkeypatch):
    if tma_min:   # bulk-copy tile kernels with one tensor-map block per destination rank (default from length 256)
        monkeypatch.setenv("SX_TMA_MIN", tma_min)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, ord_, emu_lib.path, tables, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=600)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    res = sorted(q.get(timeout=10) for _ in range(world))
    covered = 0
    for rank, err, eng, eng_ref, div, div_ref, (ista, iend) in res:
        assert err < 1e-11, (rank, err)
        assert abs(eng / eng_ref - 1) < 1e-9            # all-reduced diagnostic equals the single-rank one
        assert abs(div - div_ref) < 1e-9 * eng_ref
        covered += iend - ista + 1
    assert covered == shape[0] // 2 + 1


def _install_gloo_callbacks(p, dist, rank, world):
    import torch

    def alltoallv(send, sd, sc, recv, rd, rc):
        assert os.environ.get("SX_TEST_P2P") != "1", "the fused substep must not use the all-to-all callback with the peer-to-peer transport"
        ins = [torch.from_numpy(np.frombuffer((C.c_char * sc[r]).from_address(send + sd[r]), dtype=np.uint8).copy())
               if sc[r] else torch.empty(0, dtype=torch.uint8) for r in range(world)]
        outs = [torch.empty(rc[r], dtype=torch.uint8) for r in range(world)]
        _gloo_a2a(dist, outs, ins, rank, world)
        for r in range(world):
            if rc[r]:
                C.memmove(recv + rd[r], outs[r].numpy().ctypes.data, rc[r])

    def allreduce(ptr, n):
        t = torch.tensor([ptr[i] for i in range(n)], dtype=torch.float64)
        dist.all_reduce(t)
        for i in range(n):
            ptr[i] = float(t[i])

    p.set_comm_callbacks(alltoallv, allreduce)


def _worker_operators(rank, world, port, shape, emu_path, tables, tmpdir, q):
    """Slab-parallel stand-alone transforms (fftp.fpp:388-524, 789-919), the per-operator substep, the ROTBOUSS fused
    substep, diagnostics with MPI_MAX / MPI_SUM reductions and the field files written by several ranks."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import specter_oracle as O
    from specter_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = api.Library(emu_path)
        nx, ny, nz = shape
        p = api.Plan(nx, ny, nz, 25, 5, ord=2, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, nprocs=world, myrank=rank, lib=lib)
        _install_gloo_callbacks(p, dist, rank, world)
        g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=2)
        xs, zs = slice(p.ista - 1, p.iend), slice(p.ksta - 1, p.kend)
        nph = nz - 25
        errs = {}

        def rel(a, b):
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

        rng = np.random.default_rng(3)
        r = rng.standard_normal(g.rshape())
        dr, dc = p.real(np.ascontiguousarray(r[zs])), p.spectral()
        p.fftp3d_real_to_complex(dr, dc)
        spec = O.fftp3d_real_to_complex(g, r.copy())
        errs["r2c3"] = rel(dc.get(), spec[xs])
        p.fftp2d_real_to_complex_xy(dr, dc)
        errs["r2c2"] = rel(dc.get(), O.fftp2d_real_to_complex_xy(g, r)[xs])
        dc.put(np.ascontiguousarray(spec[xs]))
        p.fftp3d_complex_to_real(dc, dr)
        errs["c2r3"] = rel(dr.get(), O.fftp3d_complex_to_real(g, spec)[zs])
        p.fftp2d_complex_to_real_xy(dc, dr)
        errs["c2r2"] = rel(dr.get(), O.fftp2d_complex_to_real_xy(g, spec)[zs])
        # per-operator HD substep (gradre with 12 + 3 slab-parallel 3-D transforms), then output / restart files
        s = O.make_hd_state(g)
        p.hd_put_state(*[np.ascontiguousarray(a[xs]) for a in (s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)])
        p.hd_step(1e-3, 1e-3, impl=1)
        O.hd_step(g, s, 1e-3, 1e-3)
        got = p.hd_get_state()
        scale = max(np.abs(a).max() for a in (s.vx, s.vy, s.vz))
        errs["hd_modular"] = max(float(np.abs(a - b[xs]).max()) for a, b in zip(got[:3], (s.vx, s.vy, s.vz))) / scale
        v = [p.hd_field(i) for i in range(3)]
        errs["maxabs"] = abs(p.maxabs(*v, 0) / O.maxabs(g, s.vx, s.vy, s.vz, 0) - 1)
        hs = (O.energy(g, s.vx, s.vy, s.vz, 1) * O.energy(g, s.vx, s.vy, s.vz, 0)) ** 0.5
        errs["helicity"] = abs(p.helicity(*v) - O.helicity(g, s.vx, s.vy, s.vz)) / hs
        ours = os.path.join(tmpdir, "ours")
        ref = os.path.join(tmpdir, "ref")
        if rank == 0:
            os.makedirs(ours, exist_ok=True)
            os.makedirs(ref, exist_ok=True)
            O.hd_output(g, s, ref, "0001", 1e-3, outs=1)
        dist.barrier()
        p.output("HD", ours, "0001", 1e-3, outs=1)
        dist.barrier()
        worst = 0.0
        for name in ("vx", "vy", "vz", "wx", "wy", "wz", "pr"):
            a = np.fromfile(O.io_path(ours, name, "0001"))
            b = np.fromfile(O.io_path(ref, name, "0001"))
            assert a.size == b.size == nx * ny * nph, (name, a.size, b.size)
            worst = max(worst, rel(a, b) / (100 if name == "pr" else 1))
        errs["output"] = worst
        p.restart("HD", ref, "0001", 1e-3)
        want = O.hd_restart(g, ref, "0001", 1e-3)
        got = p.hd_get_state()
        scale = max(np.abs(a).max() for a in want[:3])     # global scale: the high-kx slab of a band-limited field is ~0
        errs["restart"] = max(float(np.abs(a - b[xs]).max()) for a, b in zip(got[:3], want[:3])) / scale
        # ROTBOUSS fused substep on several ranks
        b = O.make_bouss_state(g)
        p.bouss_put_state(*[np.ascontiguousarray(a[xs]) for a in (b.vx, b.vy, b.vz, b.pr, b.th, b.fx, b.fy, b.fz, b.fs)])
        om = (0.3, -0.2, 1.5)
        p.rotbouss_step(1e-3, 1e-3, 1e-3, omega=om)
        O.rotbouss_step(g, b, 1e-3, 1e-3, 1e-3, omega=om)
        got = p.bouss_get_state()
        scale = max(np.abs(a).max() for a in (b.vx, b.vy, b.vz, b.th))
        errs["rotbouss"] = max(float(np.abs(a - c[xs]).max()) for a, c in
                               zip(got[:3] + [got[4]], (b.vx, b.vy, b.vz, b.th))) / scale
        q.put((rank, errs))
        p.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (32, 16, 64)), (3, (16, 16, 64))])
def test_operators_multirank(world, shape, emu_lib, tables, tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_operators, args=(r, world, port, shape, emu_lib.path, tables, str(tmp_path), q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=900)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    for rank, errs in sorted(q.get(timeout=10) for _ in range(world)):
        for k in ("r2c3", "r2c2", "c2r3", "c2r2"):
            assert errs[k] < 1e-12, (rank, k, errs[k])
        for k in ("hd_modular", "output", "restart", "rotbouss"):
            assert errs[k] < 1e-11, (rank, k, errs[k])
        assert errs["maxabs"] < 1e-9 and errs["helicity"] < 1e-9, (rank, errs)


def _worker_solvers(rank, world, port, solver, shape, emu_path, tables, q):
    """The slab-parallel FUSED substeps of the other solvers (BOUSS, MHD, MHDBOUSS: 12 / 18 / 21 field transposes per
    substep) against the single-rank oracle; theta is compared on the physical rows of the mixed domain
    (parity_cases.phys_close)."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import specter_oracle as O
    from specter_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = api.Library(emu_path)
        nx, ny, nz = shape
        p = api.Plan(nx, ny, nz, 25, 5, ord=2, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, nprocs=world, myrank=rank, lib=lib)
        _install_gloo_callbacks(p, dist, rank, world)
        _maybe_p2p(p, dist, *{"bouss": (8, 4), "mhd": (12, 6), "mhdbouss": (14, 7)}[solver])
        g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=2)
        g.load_neumann()
        sl = slice(p.ista - 1, p.iend)
        cut = lambda arrs: [np.ascontiguousarray(a[sl]) for a in arrs]
        nph = nz - 25
        phys = lambda a: np.fft.ifft(a, axis=2)[:, :, :nph]
        if solver == "bouss":
            s = O.make_bouss_state(g)
            p.bouss_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.th, s.fx, s.fy, s.fz, s.fs)))
            p.bouss_step(1e-3, 1e-3, 1e-3)
            st = p.bouss_get_state()
            O.bouss_step(g, s, 1e-3, 1e-3, 1e-3)
            groups = [(st[:3], (s.vx, s.vy, s.vz)), ([phys(st[4])], (phys(s.th),))]
        elif solver == "mhd":
            s = O.make_mhd_state(g)
            p.mhd_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.fx, s.fy, s.fz, s.mx, s.my, s.mz)))
            p.mhd_step(1e-3, 1e-3, 5e-3, b0=(0.0, 0.0, 0.1))
            st = p.mhd_get_state()
            O.mhd_step(g, s, 1e-3, 1e-3, 5e-3, b0=(0.0, 0.0, 0.1))
            groups = [(st[:3], (s.vx, s.vy, s.vz)), (st[4:7], (s.ax, s.ay, s.az))]
        else:
            s = O.make_mhdbouss_state(g)
            p.mhdbouss_put_state(*cut((s.vx, s.vy, s.vz, s.pr, s.ax, s.ay, s.az, s.th, s.fx, s.fy, s.fz, s.mx, s.my, s.mz, s.fs)))
            p.mhdbouss_step(1e-3, 1e-3, 5e-3, 1e-3, b0=(0.0, 0.0, 0.1))
            st = p.mhdbouss_get_state()
            O.mhdbouss_step(g, s, 1e-3, 1e-3, 5e-3, 1e-3, b0=(0.0, 0.0, 0.1))
            groups = [(st[:3], (s.vx, s.vy, s.vz)), (st[4:7], (s.ax, s.ay, s.az)), ([phys(st[8])], (phys(s.th),))]
            # the <solver>_global.f90 include on several ranks: every reduction is collective, rank 0 writes the rows
            import tempfile
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import parity_cases as P
            odir = tempfile.mkdtemp()
            if world == 2:
                p.global_quantities("MHDBOUSS", odir, 2, 1e-3)
            if world != 2:
                pass
            elif rank == 0:
                rdir = tempfile.mkdtemp()
                O.solver_global(g, s, "MHDBOUSS", rdir, 2, 1e-3)
                assert sorted(os.listdir(odir)) == sorted(os.listdir(rdir))
                eng = O.energy(g, s.vx, s.vy, s.vz, 1)
                for name in os.listdir(rdir):
                    a = open(os.path.join(odir, name)).read().split("\n")[0]
                    b = open(os.path.join(rdir, name)).read().split("\n")[0]
                    assert len(a) == len(b) and a[:13] == b[:13], name
                    widths = dict(P.GLOBAL_WIDTHS, **{"balance.txt": [13, 23, 23, 23], "helicity.txt": [13, 24, 24]})[name]
                    w = 13 if "diagnostic" in name else None
                    xs = [P._fortran_float(t) for t in P._split_fixed(a, widths)[1:]]
                    ys = [P._fortran_float(t) for t in P._split_fixed(b, widths)[1:]]
                    for x, y in zip(xs, ys):
                        assert abs(x - y) <= (2e-6 if w else 1e-9) * abs(y) + 1e-9 * eng, (name, x, y)
            else:
                assert os.listdir(odir) == []
        errs = []
        for got, ref in groups:
            scale = max(np.abs(r).max() for r in ref)
            errs.append(float(max(np.abs(a - r[sl]).max() for a, r in zip(got, ref)) / scale))
        q.put((rank, errs, p.comm_stats()["exchanges"], (p.ista, p.iend)))
        p.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("solver,world,shape", [("bouss", 2, (32, 16, 64)), ("mhd", 3, (16, 16, 64)), ("mhdbouss", 2, (32, 16, 64))])
def test_fused_solvers_multirank(solver, world, shape, emu_lib, tables):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + world + {"bouss": 0, "mhd": 10, "mhdbouss": 20}[solver]
    procs = [ctx.Process(target=_worker_solvers, args=(r, world, port, solver, shape, emu_lib.path, tables, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=900)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    covered = 0
    for rank, errs, nex, (ista, iend) in sorted(q.get(timeout=10) for _ in range(world)):
        assert all(e < 1e-11 for e in errs), (solver, rank, errs)
        assert nex > 0
        covered += iend - ista + 1
    assert covered == shape[0] // 2 + 1


# ---- the peer-to-peer transport and the chunked xy pipeline (the default of the GPU build on one node) ---------------
def _worker_p2p(rank, world, port, variants, emu_path, tables, q):
    """One HD step per variant (shape, tuning environment) through the peer-to-peer transport: the receive arenas of the
    ranks are mapped into each other (POSIX shared memory here, CUDA IPC on the GPU), the blocks are copied -- or stored by
    the producing kernels -- straight into them, the xy stage runs as a pipeline over z chunks, and the callbacks carry
    nothing but the completion barrier (the all-to-all callback must stay unused)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import specter_oracle as O
    from specter_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = api.Library(emu_path)
        out = []
        for shape, env in variants:
            for k in ("SX_ZCHUNKS", "SX_TMA_MIN", "SX_P2P_DIRECT", "SX_P2P_DIRECT_PEERS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            nx, ny, nz = shape
            p = api.Plan(nx, ny, nz, 25, 5, ord=2, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, nprocs=world, myrank=rank, lib=lib)

            def alltoallv(*a):
                raise AssertionError("the all-to-all callback was used with the peer-to-peer transport")

            def allreduce(ptr, n):
                t = torch.tensor([ptr[i] for i in range(n)], dtype=torch.float64)
                dist.all_reduce(t)
                for i in range(n):
                    ptr[i] = float(t[i])

            p.set_comm_callbacks(alltoallv, allreduce)
            p.init_p2p_torch(dist, 6, 3)
            g = O.Grid(nx, ny, nz, 25, 5, Lx=1.0, Ly=0.5, Lz=1.0, tdir=tables, ord=2)
            s = O.make_hd_state(g)
            sl = slice(p.ista - 1, p.iend)
            p.hd_put_state(*[np.ascontiguousarray(a[sl]) for a in (s.vx, s.vy, s.vz, s.pr, s.fx, s.fy, s.fz)])
            errs = []
            for _ in range(2):                      # two steps: the receive arenas are reused
                p.hd_step(1e-3, 1e-3)
                O.hd_step(g, s, 1e-3, 1e-3)
                got = p.hd_get_state()
                scale = max(np.abs(a).max() for a in (s.vx, s.vy, s.vz))
                errs.append(float(max(np.abs(a - b[sl]).max() for a, b in zip(got[:3], (s.vx, s.vy, s.vz))) / scale))
            out.append((max(errs), p.comm_stats()["exchanges"], p.iend - p.ista + 1))
            p.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,variants", [
    (2, [((32, 16, 64), {}),                                               # 4 z chunks, the local block stored in place
         ((16, 128, 128), {"SX_TMA_MIN": "16", "SX_P2P_DIRECT": "2"}),     # bulk-copy tile kernels storing EVERY block into the peers
         ((32, 16, 64), {"SX_P2P_DIRECT": "0", "SX_ZCHUNKS": "8"})]),      # every block copied; more chunks than some ranks have row pairs
    (8, [((64, 16, 64), {}),                                               # the rank count of BASELINE configs[2..4]: 33 kx planes and 20 row pairs over 8 ranks
         ((32, 128, 128), {"SX_TMA_MIN": "16", "SX_P2P_DIRECT": "2"})]),   # eight tensor-map blocks per tile, every block stored into its peer
    (3, [((16, 16, 64), {"SX_ZCHUNKS": "2"}),                              # uneven slabs, 2 chunks
         ((16, 16, 64), {"SX_ZCHUNKS": "1"}),                              # no pipeline: one exchange per field
         ((32, 16, 64), {"SX_P2P_DIRECT_PEERS": "1"})]),                   # the next peer's blocks stored directly, the other copied
])
def test_fused_substep_multirank_p2p(world, variants, emu_lib, tables, monkeypatch):
    if world == 3:
        # the adversarial emulation for one of the two: random thread order, late asynchronous copies and LAZY STREAMS -- the
        # copies and the barrier of an exchange run on the communication stream only when the compute stream's wait for its
        # completion event pulls them, a kernel without that wait would read a buffer nothing has landed in
        monkeypatch.setenv("SX_EMU_ADVERSARIAL", "14")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_p2p, args=(r, world, port, variants, emu_lib.path, tables, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=900)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    res = dict(q.get(timeout=10) for _ in range(world))
    for v, (shape, env) in enumerate(variants):
        assert sum(res[r][v][2] for r in range(world)) == shape[0] // 2 + 1
        for r in range(world):
            err, nex, _ = res[r][v]
            assert err < 1e-11, (world, shape, env, r, err)
            assert nex > 0


@pytest.mark.parametrize("solver,world,shape", [("bouss", 2, (32, 16, 64)), ("mhd", 3, (16, 16, 64)), ("mhdbouss", 2, (32, 16, 64))])
def test_fused_solvers_multirank_p2p(solver, world, shape, emu_lib, tables, monkeypatch):
    monkeypatch.setenv("SX_TEST_P2P", "1")
    test_fused_solvers_multirank(solver, world, shape, emu_lib, tables)
