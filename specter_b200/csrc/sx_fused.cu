// The fused HD Runge-Kutta substep (include/hd/hd_rkstep2.f90:3-36 of the reference) as six
// memory-bound passes.  Every pass is one FFT axis with the neighbouring elementwise work and the
// slab transposition folded into its load / store pattern:
//
//   zinv_tile   v^(kz,ky,kx)      -> z-IFFT of v and of i kz v, physical rows only, written ky-fastest
//                                    in the exchange layout [rank][kx][z][ky]           (fftp.fpp:1060-1094)
//   yinv_tile   [kx][z][ky]       -> y-IFFT of (v, i ky v, dz v), written kx-fastest [z][y][kx]
//   xpass       [z][y][kx] x 9    -> 12 c2r lines (i kx applied on load), (u.grad)u / N^2, 3 r2c lines
//                                                                           (pseudospec_hd.f90:245-316)
//   yfwd_tile   [z][y][kx] x 3    -> y-FFT, written ky-fastest [kx][z][ky]
//   zfwd_rk     [kx][z][ky]       -> FC-Gram continuation + z-FFT (fftp.fpp:757-780), fc_filter
//                                    (pseudospec_hd.f90:1099-1112), Laplacian + RK update (hd_rkstep2.f90:14-32)
//   project     per (ky,kx) pencil: no-slip walls + Poisson/Laplace projection, nine z transforms in
//                                    shared memory (vboundary.f90:116-148, boundary_mod.fpp:197-402)
//
// Derivative sharing: d/dx and d/dy commute with the z-IFFT and the transposition, so only v and
// dz v (6 fields, not 12) cross the transposition; rows above the physical region never do (the
// products only use k <= pkend, pseudospec_hd.f90:255, and the forward continuation overwrites them).
//
// Thread mapping of the "tile" kernels: NP lines per CTA, lane-fastest over the NP lines
// (p = tid % NP, j = tid / NP, T = N/8 threads per line).  One side of the kernel then moves
// NP*16 B = 128 B full lines per quarter-warp and the other side 64 B pieces of NP different lines,
// which is the transposition.
#include <cstdlib>
#include <cstring>
#include "sx_fused.h"
#ifdef SX_EMU
// CPU-thread emulation (tests only): the peer-to-peer arena is a POSIX shared-memory object, so that the ranks of a
// multi-process test map each other's receive buffers the way CUDA IPC does on the GPU
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <string>
#include <unordered_map>
#include <utility>
namespace {
struct EmuShm {
  std::string name;
  size_t bytes = 0;
  std::vector<std::pair<void*, size_t>> peers;
};
std::unordered_map<const void*, EmuShm> g_emu_shm;   // keyed by the Fused object
struct EmuHandle {                                     // what travels in the 64 handle bytes
  unsigned long long bytes;
  char name[48];
};
static_assert(sizeof(EmuHandle) <= 64, "the emulated handle must fit the CUDA IPC handle");
}  // namespace
#endif

namespace sx {

// ==========================================================================================
// host side
// ==========================================================================================
static void release_fields(std::vector<cplx*>& a, const std::vector<cplx*>* alias = nullptr) {
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i] && (!alias || i >= alias->size() || (*alias)[i] != a[i])) cudaFree(a[i]);
  a.clear();
}

int fused_free(Plan& p) {
  Fused* f = p.fused;
  if (!f) return 0;
  if (f->arena) {   // R and Uz live inside the arena
    f->R.assign(f->R.size(), nullptr);
    f->Uz.assign(f->Uz.size(), nullptr);
#ifndef SX_EMU
    for (size_t r = 0; r < f->peer_arena.size(); ++r)
      if ((int)r != p.myrank && f->peer_arena[r]) cudaIpcCloseMemHandle(f->peer_arena[r]);
    cudaFree(f->arena);
#else
    EmuShm& sh = g_emu_shm[f];
    for (auto& pr : sh.peers) munmap(pr.first, pr.second);
    munmap(f->arena, sh.bytes);
    shm_unlink(sh.name.c_str());
    g_emu_shm.erase(f);
#endif
  }
  release_fields(f->R, &f->W);
  release_fields(f->W);
  release_fields(f->V);
  release_fields(f->X);
  release_fields(f->Uz, &f->U);
  release_fields(f->U);
  if (f->d_zmap) cudaFree(f->d_zmap);
  delete f;
  p.fused = nullptr;
  return 0;
}

static int fused_init(Plan& p, Fused** out) {
  if (p.fused) { *out = p.fused; return 0; }
  SX_REQUIRE(p.Cz > 0 && p.oz > 0 && p.oz <= kMaxDF, "the fused substep needs a non-periodic z direction (0 < oz <= 10)");
  Fused* f = new Fused();
  p.fused = f;
  f->nph = p.nz - p.Cz;
  zrange(f->nph, p.nprocs, p.myrank, &f->zf0, &f->nzf);
  f->nxp = (p.nxh + 7) / 8 * 8;
  f->wsize = (size_t)p.nxl * f->nph * p.ny;
  f->vsize = 0;   // set by fused_reserve (whole slab, or one z chunk)
  std::vector<ZMap> zm(f->nph);
  long long base = 0;
  for (int r = 0; r < p.nprocs; ++r) {
    int s, c;
    zrange(f->nph, p.nprocs, r, &s, &c);
    for (int q = 0; q < c; ++q) zm[s + q] = ZMap{base, c, q};
    base += (long long)p.nxl * c * p.ny;
  }
  f->z_displ.resize(p.nprocs); f->z_count.resize(p.nprocs); f->x_displ.resize(p.nprocs); f->x_count.resize(p.nprocs);
  for (int r = 0; r < p.nprocs; ++r) {
    int s, c, xs, xc;
    zrange(f->nph, p.nprocs, r, &s, &c);
    range0(p.nxh, p.nprocs, r, &xs, &xc);
    f->z_displ[r] = c ? (size_t)zm[s].base : 0;
    f->z_count[r] = (size_t)p.nxl * c * p.ny;
    f->x_displ[r] = (size_t)xs * f->nzf * p.ny;
    f->x_count[r] = (size_t)xc * f->nzf * p.ny;
  }
  SX_CUDA_CHECK(cudaMalloc((void**)&f->d_zmap, zm.size() * sizeof(ZMap)));
  SX_CUDA_CHECK(cudaMemcpy(f->d_zmap, zm.data(), zm.size() * sizeof(ZMap), cudaMemcpyHostToDevice));
  *out = f;
  return 0;
}

static bool chunked(const Plan& p, const Fused& f);

// grow the work-field pools: nw transposed inverse fields, nv real-side inputs of the x pass, nx nonlinear terms.
// window: the caller runs the z-chunked xy pipeline, so V / X only need the rows of one chunk (Fused::vwin).
static int fused_reserve(Plan& p, Fused& f, int nw, int nv, int nx, bool window) {
  const size_t rsize = (size_t)p.nxh * f.nzf * p.ny;  // y-stage side [kx][zl][ky]
  const size_t both = f.wsize > rsize ? f.wsize : rsize;
  if (f.arena) SX_REQUIRE(nw <= f.arena_nw && nx <= f.arena_nx, "the peer-to-peer arena was exported for fewer fields than this solver transposes");
  while ((int)f.W.size() < nw) {
    cplx *w = nullptr, *r = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&w, both * sizeof(cplx)));
    if (p.nprocs == 1) r = w;
    else if (f.arena) r = f.arena + f.W.size() * arena_rs(p, p.myrank);
    else SX_CUDA_CHECK(cudaMalloc((void**)&r, rsize * sizeof(cplx)));
    f.W.push_back(w);
    f.R.push_back(r);
  }
  const int rows = window ? cdiv(f.nzf, p.knob_zchunks) : f.nzf;
  f.vwin = window && rows < f.nzf;
  if (rows > f.vrows) {   // a solver that needs whole-slab V / X after a windowed one: start the pools again
    SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
    release_fields(f.V);
    release_fields(f.X);
    f.vrows = rows;
    f.vsize = (size_t)rows * p.ny * f.nxp;
  }
  while ((int)f.V.size() < nv) {
    cplx* v = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&v, (f.vsize ? f.vsize : 1) * sizeof(cplx)));
    f.V.push_back(v);
  }
  while ((int)f.X.size() < nx) {
    cplx* x = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&x, (f.vsize ? f.vsize : 1) * sizeof(cplx)));
    f.X.push_back(x);
  }
  while ((int)f.U.size() < nx) {
    cplx *u = nullptr, *uz = nullptr;
    SX_CUDA_CHECK(cudaMalloc((void**)&u, both * sizeof(cplx)));
    if (p.nprocs == 1) uz = u;
    else if (f.arena) uz = f.arena + f.arena_nw * arena_rs(p, p.myrank) + f.U.size() * arena_ws(p, p.myrank);
    else SX_CUDA_CHECK(cudaMalloc((void**)&uz, f.wsize * sizeof(cplx)));
    f.U.push_back(u);
    f.Uz.push_back(uz);
  }
  return 0;
}

// W[slot] (z side, [rank][kxl][zl][ky]) -> R[slot] (xy side, [kx][zl][ky]); event slots 0..15
static int to_real_begin(Plan& p, Fused& f, int slot) {
  if (p.nprocs == 1) return 0;
  if (f.p2p) {
    std::vector<cplx*> dst(p.nprocs);
    std::vector<size_t> cnt(f.z_count);
    for (int r = 0; r < p.nprocs; ++r) {
      dst[r] = peer_r_dst(p, f, slot, r);
      // blocks the z-inverse kernel already stored into their destination (sx_fused_tiles.cu) are not copied
      if (f.zinv_direct && direct_to(p, f, r)) cnt[r] = 0;
    }
    return exchange_begin_p2p(p, slot, f.W[slot], f.z_displ.data(), cnt.data(), dst.data());
  }
  return exchange_begin(p, slot, f.W[slot], f.R[slot], f.z_displ.data(), f.z_count.data(), f.x_displ.data(), f.x_count.data());
}
// U[slot] (xy side) -> Uz[slot] (z side); event slots 16..23
static int to_spec_begin(Plan& p, Fused& f, int slot) {
  if (p.nprocs == 1) return 0;
  if (f.p2p) {
    std::vector<cplx*> dst(p.nprocs);
    std::vector<size_t> cnt(f.x_count);
    for (int r = 0; r < p.nprocs; ++r) {
      dst[r] = peer_uz_dst(p, f, slot, r);
      if (f.yfwd_direct && direct_to(p, f, r)) cnt[r] = 0;
    }
    return exchange_begin_p2p(p, 16 + slot, f.U[slot], f.x_displ.data(), cnt.data(), dst.data());
  }
  return exchange_begin(p, 16 + slot, f.U[slot], f.Uz[slot], f.x_displ.data(), f.x_count.data(), f.z_displ.data(), f.z_count.data());
}
static int ex_wait(Plan& p, int ev) {
  if (p.nprocs == 1) return 0;
  // chunked pipeline: the inverse side was waited for chunk by chunk; the way back of component c is complete when the
  // per-component round of the last chunk (event slot 24 + c) is -- the rounds complete in issue order
  if (p.fused && p.fused->chunked_now) return ev >= 16 ? exchange_wait(p, 24 + (ev - 16)) : 0;
  return exchange_wait(p, ev);
}

static int fused_begin(Plan& p, Fused** fp, int nw, int nv, int nx, bool may_chunk = false) {
  if (fused_init(p, fp)) return 1;
  SX_REQUIRE(comm_ready(p), "multi-rank plan without a communicator: call sx_plan_set_comm or sx_plan_set_comm_callbacks");
  return fused_reserve(p, **fp, nw, nv, nx, may_chunk && chunked(p, **fp));
}

// host-buffer entries (sx_hd_step_host): the upload of state field i is still in flight on the copy stream; the first
// kernel that reads it waits for its event (once)
static int consume_wait(Plan& p, int i) {
  if (p.pre_wait[i]) {
    SX_CUDA_CHECK(cudaStreamWaitEvent(p.stream, p.pre_wait[i], 0));
    p.pre_wait[i] = nullptr;
  }
  return 0;
}

// inverse half shared by HD and BOUSS: q_c -> V[c], V[NC+c] (dy), V[2NC+c] (dz)
template <int NC> static int gradient_fields_to_real(Plan& p, Fused& f, const cplx* const* q) {
  for (int c = 0; c < NC; ++c) {
    if (c < 3 && consume_wait(p, c)) return 1;
    if (fused_zinv(p, f, q[c], f.W[2 * c], f.W[2 * c + 1])) return 1;
    if (to_real_begin(p, f, 2 * c) || to_real_begin(p, f, 2 * c + 1)) return 1;
  }
  for (int c = 0; c < NC; ++c) {
    if (ex_wait(p, 2 * c) || ex_wait(p, 2 * c + 1)) return 1;
    if (fused_yinv(p, f, f.R[2 * c], f.V[c], f.V[NC + c])) return 1;
    if (fused_yinv(p, f, f.R[2 * c + 1], f.V[2 * NC + c], nullptr)) return 1;
  }
  return 0;
}
// ------------------------------------------------------------------------------------------
// Multi-rank xy stage as a pipeline over z chunks (peer-to-peer transport only).  The exchanges of the 2*NC
// inverse fields and of the NC nonlinear terms sit on the critical path on both sides of the x pass; cut into
// NCH z chunks, chunk k of every field travels while chunk k-1 is in the y-inverse / x pass / y-forward
// kernels, and the way back of chunk k travels under the kernels of chunk k+1:
//   zinv(all fields) -> [in-round k: rows of chunk k of every field into every rank's R, barrier] k = 0..NCH-1
//   for k: wait in-round k; yinv, xpass, yfwd on the z window of chunk k; out-round k (rows of chunk k of U -> Uz)
// All rounds run on the communication stream in issue order; the chain of barriers orders buffer reuse as before.
// ------------------------------------------------------------------------------------------
static bool chunked(const Plan& p, const Fused& f) {
  return p.nprocs > 1 && f.p2p && p.knob_zchunks > 1 && p.knob_zchunks <= 8;
}


template <int NC> static int xy_stage_chunked(Plan& p, Fused& f, const cplx* const* q) {
  const int nch = p.knob_zchunks;
  int xs_me, xc_me;
  range0(p.nxh, p.nprocs, p.myrank, &xs_me, &xc_me);
  std::vector<int> zc(p.nprocs);
  for (int r = 0; r < p.nprocs; ++r) {
    int s;
    zrange(f.nph, p.nprocs, r, &s, &zc[r]);
  }
  // z stage: all inverse fields (the local block goes straight into R when the tensor-map kernel runs)
  for (int c = 0; c < NC; ++c) {
    if (c < 3 && consume_wait(p, c)) return 1;
    if (fused_zinv(p, f, q[c], f.W[2 * c], f.W[2 * c + 1])) return 1;
    if (p2p_mark(p, c)) return 1;
  }
  const bool zdirect = f.zinv_direct && f.direct >= 1;
  std::vector<P2PCopy> cp;
  for (int k = 0; k < nch; ++k) {
    for (int c = 0; c < NC; ++c) {   // sub-round per component: starts as soon as that component's zinv is done
      cp.clear();
      for (int h = 0; h < 2; ++h)
        for (int qd = 1; qd <= p.nprocs; ++qd) {
          const int r = (p.myrank + qd) % p.nprocs;
          if (zdirect && direct_to(p, f, r)) continue;
          int c0, cc;
          range0(zc[r], nch, k, &c0, &cc);
          cp.push_back(P2PCopy{peer_r_dst(p, f, 2 * c + h, r) + (size_t)c0 * p.ny, f.W[2 * c + h] + f.z_displ[r] + (size_t)c0 * p.ny,
                               (size_t)cc * p.ny, (size_t)p.nxl, (size_t)zc[r] * p.ny, (size_t)zc[r] * p.ny, r != p.myrank});
        }
      const int w = c;
      if (p2p_round(p, &w, k == 0 ? 1 : 0, cp.data(), (int)cp.size(), c == NC - 1, c == NC - 1 ? k : -1)) return 1;
    }
  }
  f.chunked_now = true;
  for (int k = 0; k < nch; ++k) {
    int c0, cc;
    range0(f.nzf, nch, k, &c0, &cc);
    if (exchange_wait(p, k)) return 1;
    f.zw0 = c0;
    f.zwc = cc;
    for (int c = 0; c < NC; ++c) {
      if (fused_yinv(p, f, f.R[2 * c], f.V[c], f.V[NC + c])) return 1;
      if (fused_yinv(p, f, f.R[2 * c + 1], f.V[2 * NC + c], nullptr)) return 1;
    }
    if (fused_xpass(p, f, NC, p.d_kxg)) return 1;
    auto out_copies = [&](int c) {   // after the y-forward launch of component c (its launcher sets yfwd_direct)
      const bool ydirect = f.yfwd_direct && f.direct >= 1;
      for (int qd = 1; qd <= p.nprocs; ++qd) {
        const int r = (p.myrank + qd) % p.nprocs;
        if (ydirect && direct_to(p, f, r)) continue;
        int xs, xc;
        range0(p.nxh, p.nprocs, r, &xs, &xc);
        cp.push_back(P2PCopy{peer_uz_dst(p, f, c, r) + (size_t)c0 * p.ny, f.U[c] + f.x_displ[r] + (size_t)c0 * p.ny,
                             (size_t)cc * p.ny, (size_t)xc, (size_t)f.nzf * p.ny, (size_t)f.nzf * p.ny, r != p.myrank});
      }
    };
    if (k < nch - 1) {
      for (int c = 0; c < NC; ++c)
        if (fused_yfwd(p, f, f.X[c], f.U[c])) return 1;
      if (p2p_mark(p, 16 + k)) return 1;
      cp.clear();
      for (int c = 0; c < NC; ++c) out_copies(c);
      const int w = 16 + k;
      if (p2p_round(p, &w, 1, cp.data(), (int)cp.size(), true, 16 + k)) return 1;
    } else {
      // last chunk: one round per component, in the order the z stage consumes them (theta first in BOUSS), so that
      // the z-forward kernel of a component starts when ITS last rows have landed while the others still travel
      for (int i = 0; i < NC; ++i) {
        const int c = NC == 4 ? (i + 3) % 4 : i;
        if (fused_yfwd(p, f, f.X[c], f.U[c])) return 1;
        if (p2p_mark(p, 24 + c)) return 1;
        cp.clear();
        out_copies(c);
        const int w = 24 + c;
        if (p2p_round(p, &w, 1, cp.data(), (int)cp.size(), true, 24 + c)) return 1;
      }
    }
  }
  f.zwc = -1;
  f.zw0 = 0;
  return 0;   // the z stage waits per component (ex_wait: event slots 24 + c)
}

static int nonlinear_to_spectral_begin(Plan& p, Fused& f, int nx) {
  for (int c = 0; c < nx; ++c) {
    if (fused_yfwd(p, f, f.X[c], f.U[c])) return 1;
    if (to_spec_begin(p, f, c)) return 1;
  }
  return 0;
}

int fused_p2p_export(Plan& p, int nw, int nx, void* handle64) {
#ifndef SX_EMU
  SX_REQUIRE(p.nprocs > 1, "sx_plan_p2p_export: single-rank plan");
  Fused* fp;
  if (fused_init(p, &fp)) return 1;
  Fused& f = *fp;
  SX_REQUIRE(f.arena == nullptr && f.W.empty() && f.X.empty(), "sx_plan_p2p_export: call once, before the first substep");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are expected to be 64 bytes");
  const size_t elems = (size_t)nw * arena_rs(p, p.myrank) + (size_t)nx * arena_ws(p, p.myrank);
  SX_CUDA_CHECK(cudaMalloc((void**)&f.arena, elems * sizeof(cplx)));
  f.arena_nw = nw;
  f.arena_nx = nx;
  cudaIpcMemHandle_t h;
  SX_CUDA_CHECK(cudaIpcGetMemHandle(&h, f.arena));
  memcpy(handle64, &h, sizeof(h));
  return 0;
#else
  SX_REQUIRE(p.nprocs > 1, "sx_plan_p2p_export: single-rank plan");
  Fused* fp;
  if (fused_init(p, &fp)) return 1;
  Fused& f = *fp;
  SX_REQUIRE(f.arena == nullptr && f.W.empty() && f.X.empty(), "sx_plan_p2p_export: call once, before the first substep");
  const size_t elems = (size_t)nw * arena_rs(p, p.myrank) + (size_t)nx * arena_ws(p, p.myrank);
  static int counter = 0;
  EmuShm sh;
  sh.name = "/sxemu." + std::to_string((long)getpid()) + "." + std::to_string(counter++);
  sh.bytes = elems * sizeof(cplx);
  const int fd = shm_open(sh.name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  SX_REQUIRE(fd >= 0 && ftruncate(fd, (off_t)sh.bytes) == 0, "emulated arena: shm_open / ftruncate failed");
  void* ptr = mmap(nullptr, sh.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  SX_REQUIRE(ptr != MAP_FAILED, "emulated arena: mmap failed");
  f.arena = (cplx*)ptr;
  f.arena_nw = nw;
  f.arena_nx = nx;
  EmuHandle h;
  memset(&h, 0, sizeof(h));
  h.bytes = sh.bytes;
  SX_REQUIRE(sh.name.size() < sizeof(h.name), "emulated arena: name too long");
  memcpy(h.name, sh.name.c_str(), sh.name.size());
  memset(handle64, 0, 64);
  memcpy(handle64, &h, sizeof(h));
  g_emu_shm[&f] = sh;
  return 0;
#endif
}

int fused_p2p_import(Plan& p, const void* handles) {
#ifndef SX_EMU
  Fused* f = p.fused;
  SX_REQUIRE(f != nullptr && f->arena != nullptr, "sx_plan_p2p_import: call sx_plan_p2p_export first");
  f->peer_arena.assign(p.nprocs, nullptr);
  for (int r = 0; r < p.nprocs; ++r) {
    if (r == p.myrank) { f->peer_arena[r] = f->arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    SX_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    f->peer_arena[r] = (cplx*)ptr;
  }
  f->p2p = true;
  if (const char* e = getenv("SX_P2P_DIRECT")) {
    SX_REQUIRE(e[0] >= '0' && e[0] <= '2' && e[1] == 0, "invalid value of the tuning variable SX_P2P_DIRECT (0, 1 or 2)");
    f->direct = e[0] - '0';
  }
  if (const char* e = getenv("SX_P2P_DIRECT_PEERS")) {
    SX_REQUIRE(e[0] >= '0' && e[0] <= '7' && e[1] == 0, "invalid value of the tuning variable SX_P2P_DIRECT_PEERS (0..7)");
    f->direct_peers = e[0] - '0';
  }
  return 0;
#else
  Fused* f = p.fused;
  SX_REQUIRE(f != nullptr && f->arena != nullptr, "sx_plan_p2p_import: call sx_plan_p2p_export first");
  f->peer_arena.assign(p.nprocs, nullptr);
  EmuShm& sh = g_emu_shm[f];
  for (int r = 0; r < p.nprocs; ++r) {
    if (r == p.myrank) { f->peer_arena[r] = f->arena; continue; }
    EmuHandle h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, sizeof(h));
    h.name[sizeof(h.name) - 1] = 0;
    const int fd = shm_open(h.name, O_RDWR, 0600);
    SX_REQUIRE(fd >= 0, "emulated arena: cannot open the arena of a peer rank");
    void* ptr = mmap(nullptr, (size_t)h.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    SX_REQUIRE(ptr != MAP_FAILED, "emulated arena: mmap of a peer arena failed");
    sh.peers.emplace_back(ptr, (size_t)h.bytes);
    f->peer_arena[r] = (cplx*)ptr;
  }
  f->p2p = true;
  if (const char* e = getenv("SX_P2P_DIRECT")) {
    SX_REQUIRE(e[0] >= '0' && e[0] <= '2' && e[1] == 0, "invalid value of the tuning variable SX_P2P_DIRECT (0, 1 or 2)");
    f->direct = e[0] - '0';
  }
  if (const char* e = getenv("SX_P2P_DIRECT_PEERS")) {
    SX_REQUIRE(e[0] >= '0' && e[0] <= '7' && e[1] == 0, "invalid value of the tuning variable SX_P2P_DIRECT_PEERS (0..7)");
    f->direct_peers = e[0] - '0';
  }
  return 0;
#endif
}

// velocity part of every solver's substep: nonlinear terms Uz[0..2] -> RK update of st[0..2] (base st[7..9], forcing
// st[4..6]) -> v_imposebc_and_project.  Three z-forward / RK launches (memory-bound, 0.82 of the HBM peak) and the
// projection kernel: a single merged kernel was built and measured in round 2 and LOST (7.4 ms against 5.8 ms: the
// per-pencil state of the projection forces 8 warps per SM onto the streaming RK part as well;
// profiles/r2_zstage_experiment.md), so the two stay separate.
static int velocity_zstage(Plan& p, Fused& f, cplx* const* st, const RkTerm* rk, int o, double dt, double rmp,
                           const double* zs, const double* ze) {
  for (int c = 0; c < 3; ++c) {
    if (ex_wait(p, 16 + c)) return 1;
    if (fused_zfwd_rk(p, f, f.Uz[c], st[c], st[c], st[7 + c], st[4 + c], rk[c], dt, rmp)) return 1;
  }
  if (consume_wait(p, 3)) return 1;
  return fused_project(p, f, st[0], st[1], st[2], st[3], o, zs, ze);
}

// hd_rkstep2.f90:3-36.  st[0..2] v, st[3] pr, st[4..6] f, st[7..9] RK base.
int hd_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, const double* zs, const double* ze) {
  Fused* fp;
  if (fused_begin(p, &fp, 6, 9, 3, true)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  f.chunked_now = false;
  if (chunked(p, f)) {
    if (xy_stage_chunked<3>(p, f, st)) return 1;
  } else {
    if (gradient_fields_to_real<3>(p, f, st)) return 1;
    if (fused_xpass(p, f, 3, p.d_kxg)) return 1;
    if (nonlinear_to_spectral_begin(p, f, 3)) return 1;
  }
  RkTerm rk[3];
  for (int c = 0; c < 3; ++c) rk[c].cL = nu;
  return velocity_zstage(p, f, st, rk, o, dt, rmp, zs, ze);
}

int mhd_curls(Plan& p, const cplx* vx, const cplx* vy, const cplx* vz, cplx* ax, cplx* ay, cplx* az, cplx* const* W,
              cplx* const* B, const double* b0);
static int a_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph);
int s_imposebc(Plan& p, cplx* th);
int theta_roundtrip(Plan& p, cplx* th, cplx* out);
int a_imposebc_and_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph);

// bouss_rkstep2.f90:3-59.  st[0..2] v, 3 pr, 4..6 f, 7..9 C1..C3, 10 th, 11 fs, 12 C7, 13 scratch.
int bouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                        const double* zs, const double* ze) {
  Fused* fp;
  if (fused_begin(p, &fp, 8, 12, 4, true)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  const cplx* q[4] = {st[0], st[1], st[2], st[10]};
  f.chunked_now = false;
  if (chunked(p, f)) {
    if (xy_stage_chunked<4>(p, f, q)) return 1;
  } else {
    if (gradient_fields_to_real<4>(p, f, q)) return 1;
    if (fused_xpass(p, f, 4, p.d_kxg)) return 1;          // gradre (3) and advect (1) in one pass
    if (nonlinear_to_spectral_begin(p, f, 4)) return 1;
  }
  // theta first, into the scratch field: it reads the not yet updated v_z (heat current), and v_z below reads
  // the not yet updated theta (buoyancy)
  RkTerm rt;
  rt.cL = kappa; rt.couple = st[2]; rt.ccoef = -xtemp;
  if (ex_wait(p, 16 + 3)) return 1;
  if (fused_zfwd_rk(p, f, f.Uz[3], st[10], st[13], st[12], st[11], rt, dt, rmp)) return 1;
  {
    RkTerm rk[3];
    for (int c = 0; c < 3; ++c) rk[c].cL = nu;
    rk[2].couple = st[10];
    rk[2].ccoef = -xmom;
    if (velocity_zstage(p, f, st, rk, o, dt, rmp, zs, ze)) return 1;
  }
  // s_imposebc, fc_filter and the theta round trip (bouss_rkstep2.f90:53-59); all pencil-local
  return s_imposebc(p, st[13]) || op_fc_filter(p, st[13]) || theta_roundtrip(p, st[13], st[10]);
}

int rot_couple(Plan& p, cplx* const* f, const double* om, double xmom, cplx* cx, cplx* cy, cplx* cz);

// rotbouss_rkstep2.f90:3-56 on the BOUSS state: the fused BOUSS passes with the Coriolis (+ buoyancy) terms handed
// to the z-forward / RK kernel as one precomputed coupling field per component (they read the not yet updated v, th),
// and theta updated in place (no filter / round trip at the end).
int rotbouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double kappa, double xmom,
                           double xtemp, const double* om, const double* zs, const double* ze) {
  Fused* fp;
  if (fused_begin(p, &fp, 8, 12, 4, true)) return 1;
  Fused& f = *fp;
  const double rmp = 1.0 / (double)o;
  cplx* cpl[3];
  for (int c = 0; c < 3; ++c) if (plan_cwork(p, 9 + c, &cpl[c])) return 1;
  if (rot_couple(p, st, om, xmom, cpl[0], cpl[1], cpl[2])) return 1;
  const cplx* q[4] = {st[0], st[1], st[2], st[10]};
  f.chunked_now = false;
  if (chunked(p, f)) {
    if (xy_stage_chunked<4>(p, f, q)) return 1;
  } else {
    if (gradient_fields_to_real<4>(p, f, q)) return 1;
    if (fused_xpass(p, f, 4, p.d_kxg)) return 1;
    if (nonlinear_to_spectral_begin(p, f, 4)) return 1;
  }
  RkTerm rt;   // theta first: the heat current reads the not yet updated v_z
  rt.cL = kappa; rt.couple = st[2]; rt.ccoef = -xtemp;
  if (ex_wait(p, 16 + 3)) return 1;
  if (fused_zfwd_rk(p, f, f.Uz[3], st[10], st[10], st[12], st[11], rt, dt, rmp)) return 1;
  {
    RkTerm rk[3];
    for (int c = 0; c < 3; ++c) { rk[c].cL = nu; rk[c].couple = cpl[c]; rk[c].ccoef = 1.0; }
    if (velocity_zstage(p, f, st, rk, o, dt, rmp, zs, ze)) return 1;
  }
  return s_imposebc(p, st[10]);
}

// vector-potential boundary step of the fused MHD substeps: one pencil kernel where it applies
static int a_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph) {
  const int rc = fused_aproject(p, ax, ay, az, ph);
  return rc >= 0 ? rc : a_imposebc_and_project(p, ax, ay, az, ph);
}

// mhd_rkstep2.f90:3-84.  st[0..2] v, 3 pr, 4..6 f, 7..9 C1..C3, 10..12 a, 13 ph, 14..16 m, 17..19 C9..C11.
int mhd_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double mu, const double* b0) {
  Fused* fp;
  if (fused_begin(p, &fp, 12, 12, 6)) return 1;
  Fused& f = *fp;
  f.chunked_now = false;
  const double rmp = 1.0 / (double)o;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  cplx *B[3], *Wv[3];
  for (int c = 0; c < 3; ++c)
    if (plan_cwork(p, 9 + c, &B[c]) || plan_cwork(p, 12 + c, &Wv[c])) return 1;
  cplx *ax = st[10], *ay = st[11], *az = st[12];
  // B = curl A (+ uniform field at the mean mode), J = curl B written over A, omega = curl v   (:6-20, prodre): one pass
  if (mhd_curls(p, st[0], st[1], st[2], ax, ay, az, Wv, B, b0)) return 1;
  // twelve plain fields to real space: v -> V[0..2], omega -> V[3..5], B -> V[6..8], J -> V[9..11]
  const cplx* q[12] = {st[0], st[1], st[2], Wv[0], Wv[1], Wv[2], B[0], B[1], B[2], ax, ay, az};
  for (int c = 0; c < 12; ++c) {
    if (fused_zinv(p, f, q[c], f.W[c], nullptr)) return 1;
    if (to_real_begin(p, f, c)) return 1;
  }
  for (int c = 0; c < 12; ++c) {
    if (ex_wait(p, c)) return 1;
    if (fused_yinv(p, f, f.R[c], f.V[c], nullptr)) return 1;
  }
  // X[0..2] = omega x v - J x B = -(v x omega) + (B x J);  X[3..5] = v x B
  {
    const int Pi[2] = {0, 6}, Qi[2] = {3, 9};
    const double sg[2] = {-1.0, 1.0};
    if (fused_xcross(p, f, 2, Pi, Qi, sg, 0)) return 1;
    const int Pe[1] = {0}, Qe[1] = {6};
    const double se[1] = {1.0};
    if (fused_xcross(p, f, 1, Pe, Qe, se, 3)) return 1;
  }
  if (nonlinear_to_spectral_begin(p, f, 6)) return 1;
  {
    RkTerm rk[3];
    for (int c = 0; c < 3; ++c) rk[c].cL = nu;
    if (velocity_zstage(p, f, st, rk, o, dt, rmp, nullptr, nullptr)) return 1;
  }
  for (int c = 0; c < 3; ++c) {
    RkTerm rk;
    rk.cL = -mu; rk.lap = 0; rk.sNL = 1.0;
    if (ex_wait(p, 16 + 3 + c)) return 1;
    if (fused_zfwd_rk(p, f, f.Uz[3 + c], st[10 + c], st[10 + c], st[17 + c], st[14 + c], rk, dt, rmp)) return 1;
  }
  return a_project(p, ax, ay, az, st[13]);
}

// mhdbouss_rkstep2.f90:3-106.  st: the MHD slots (0..19), 20 th, 21 fs, 22 C7.  The MHD passes plus theta: its
// z-inverse (theta, dz theta), y-inverse (theta, dy theta; dz theta), a scalar-advection x pass fed by the velocity
// lines the cross-product pass already reads, and a seventh nonlinear term on the way back.  theta is updated into
// a scratch field first (the heat current reads the not yet updated v_z, the buoyancy the not yet updated theta) and
// lands in st[20] through s_imposebc and the round trip of :102-105 (no filter here, unlike BOUSS).
int mhdbouss_rkstep2_fused(Plan& p, cplx* const* st, int o, double dt, double nu, double mu, double kappa, double xmom,
                           double xtemp, const double* b0) {
  Fused* fp;
  if (fused_begin(p, &fp, 14, 15, 7)) return 1;
  Fused& f = *fp;
  f.chunked_now = false;
  const double rmp = 1.0 / (double)o;
  const double N = (double)p.nx * (double)p.ny * (double)p.nz;
  cplx *B[3], *Wv[3], *thn;
  for (int c = 0; c < 3; ++c)
    if (plan_cwork(p, 9 + c, &B[c]) || plan_cwork(p, 12 + c, &Wv[c])) return 1;
  if (plan_cwork(p, 15, &thn)) return 1;
  cplx *ax = st[10], *ay = st[11], *az = st[12], *th = st[20];
  if (mhd_curls(p, st[0], st[1], st[2], ax, ay, az, Wv, B, b0)) return 1;
  const cplx* q[12] = {st[0], st[1], st[2], Wv[0], Wv[1], Wv[2], B[0], B[1], B[2], ax, ay, az};
  for (int c = 0; c < 12; ++c) {
    if (fused_zinv(p, f, q[c], f.W[c], nullptr)) return 1;
    if (to_real_begin(p, f, c)) return 1;
  }
  if (fused_zinv(p, f, th, f.W[12], f.W[13])) return 1;
  if (to_real_begin(p, f, 12) || to_real_begin(p, f, 13)) return 1;
  for (int c = 0; c < 12; ++c) {
    if (ex_wait(p, c)) return 1;
    if (fused_yinv(p, f, f.R[c], f.V[c], nullptr)) return 1;
  }
  if (ex_wait(p, 12) || ex_wait(p, 13)) return 1;
  if (fused_yinv(p, f, f.R[12], f.V[12], f.V[13])) return 1;     // theta, dy theta
  if (fused_yinv(p, f, f.R[13], f.V[14], nullptr)) return 1;     // dz theta
  {
    const int Pi[2] = {0, 6}, Qi[2] = {3, 9};
    const double sg[2] = {-1.0, 1.0};
    if (fused_xcross(p, f, 2, Pi, Qi, sg, 0)) return 1;            // omega x v - J x B
    const int Pe[1] = {0}, Qe[1] = {6};
    const double se[1] = {1.0};
    if (fused_xcross(p, f, 1, Pe, Qe, se, 3)) return 1;            // v x B
  }
  if (fused_xadvect(p, f, 0, 12, 6, p.d_kxg)) return 1;            // v . grad theta
  if (nonlinear_to_spectral_begin(p, f, 7)) return 1;
  RkTerm rt;
  rt.cL = kappa; rt.couple = st[2]; rt.ccoef = -xtemp;
  if (ex_wait(p, 16 + 6)) return 1;
  if (fused_zfwd_rk(p, f, f.Uz[6], th, thn, st[22], st[21], rt, dt, rmp)) return 1;
  {
    RkTerm rk[3];
    for (int c = 0; c < 3; ++c) rk[c].cL = nu;
    rk[2].couple = th;
    rk[2].ccoef = -xmom;
    if (velocity_zstage(p, f, st, rk, o, dt, rmp, nullptr, nullptr)) return 1;
  }
  for (int c = 0; c < 3; ++c) {
    RkTerm rk;
    rk.cL = -mu; rk.lap = 0; rk.sNL = 1.0;
    if (ex_wait(p, 16 + 3 + c)) return 1;
    if (fused_zfwd_rk(p, f, f.Uz[3 + c], st[10 + c], st[10 + c], st[17 + c], st[14 + c], rk, dt, rmp)) return 1;
  }
  if (a_project(p, ax, ay, az, st[13])) return 1;
  return s_imposebc(p, thn) || theta_roundtrip(p, thn, th);
}

}  // namespace sx
