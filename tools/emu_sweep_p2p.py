"""Randomised multi-rank sweep on the CPU-thread emulation: the peer-to-peer pipeline on 2 / 3 / 4 / 5 / 8 rank processes (gloo), random shapes,
z-chunk counts and direct-store settings, two HD steps each against the single-rank oracle at 1e-11.   python tools/emu_sweep_p2p.py"""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch.multiprocessing as mp
import test_multirank_gloo as G
from specter_b200 import build
emu = build.build_emu()
T = os.path.join(ROOT, 'tests', 'golden', 'tables')
random.seed(7)
shapes=[(32,16,64),(16,32,64),(64,16,128),(16,16,64),(32,32,64)]
def variants(n):
    out=[]
    for _ in range(n):
        env={}
        z=random.choice(["1","2","3","4","8"]); env["SX_ZCHUNKS"]=z
        d=random.choice(["0","1","2"]); env["SX_P2P_DIRECT"]=d
        if random.random()<0.4: env["SX_P2P_DIRECT_PEERS"]=random.choice(["1","2","3"])
        if random.random()<0.4: env["SX_TMA_MIN"]="16"
        out.append((random.choice(shapes), env))
    return out
if __name__ == "__main__":
    fails=0
    for world in (2,3,4,5,8):
        vs=variants(8)
        ctx=mp.get_context("spawn"); q=ctx.Queue(); port=36500+world
        procs=[ctx.Process(target=G._worker_p2p, args=(r,world,port,vs,emu,T,q)) for r in range(world)]
        [p.start() for p in procs]; [p.join(timeout=1500) for p in procs]
        codes=[p.exitcode for p in procs]
        if any(c!=0 for c in codes): print("world",world,"exit codes",codes); fails+=1; continue
        res=dict(q.get(timeout=10) for _ in range(world))
        for v,(shape,env) in enumerate(vs):
            worst=max(res[r][v][0] for r in range(world))
            ok = worst<1e-11 and sum(res[r][v][2] for r in range(world))==shape[0]//2+1
            if not ok: fails+=1
            print("world",world,shape,env,"err %.1e"%worst,"ok" if ok else "FAIL", flush=True)
    print("failures:",fails)
