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
