// Slab exchange between the z-stage (kx-slabs, all z) and the xy-stage (z-slabs, all kx), and the
// scalar reductions of the diagnostics.  Replaces the reference's ring of MPI_ISEND/MPI_IRECV with
// derived datatypes (fftp/fftp.fpp:478-499, 670-691, 887-907, 1024-1044) and its MPI_REDUCE calls
// (e.g. pseudospec_hd.f90:631).
//
// The FFT kernels on both sides already write / read the per-destination blocks contiguously
// ([rank][kxl][zl][ky] on the z side, [kx][zl][ky] on the xy side), so the exchange is a plain
// all-to-all-v of contiguous blocks: one ncclGroup of ncclSend/ncclRecv on a dedicated stream,
// ordered against the compute stream with events so that the exchange of field c overlaps the
// transforms of field c+1.  No packing kernels, no host staging.
//
// A caller may instead install callbacks (sx_plan_set_comm_callbacks): the Fortran/MPI driver can
// route the blocks through CUDA-aware MPI_Alltoallv, and the CPU test-suite routes them through
// torch.distributed/gloo under the kernel emulation.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/specter_b200.h"
#include "sx_plan.h"
#ifndef SX_EMU
#include <nccl.h>
#endif

namespace sx {

struct Comm {
#ifndef SX_EMU
  ncclComm_t nccl = nullptr;
#endif
  cudaStream_t stream = nullptr;
  sx_alltoallv_fn a2a = nullptr;
  sx_allreduce_fn allred = nullptr;
  void* user = nullptr;
  cudaEvent_t ready[32] = {nullptr}, done[32] = {nullptr};
  // timing mode: one event pair per exchange, resolved lazily (comm_resolve) so that the exchange is never
  // synchronised with the host and keeps overlapping the compute stream while it is being measured
  std::vector<cudaEvent_t> tev;
  size_t ntev = 0;
  double* d_scal = nullptr;
  double* h_scal = nullptr;
  float* d_bar = nullptr;   // payload of the completion barrier of the peer-to-peer exchange
  static constexpr int kSide = 7;
  cudaStream_t side[kSide] = {nullptr};   // extra copy streams: one copy engine does not fill NVLink
  cudaEvent_t fork = nullptr, join[kSide] = {nullptr};
  int nsplit = 2;
  // accounting for the NVLink roofline
  double bytes_sent = 0.0, ms = 0.0;
  long long exchanges = 0;
};

int comm_free(Plan& p) {
  Comm* c = p.comm;
  if (!c) return 0;
#ifndef SX_EMU
  if (c->nccl) ncclCommDestroy(c->nccl);
#endif
  for (int i = 0; i < 32; ++i) {
    if (c->ready[i]) cudaEventDestroy(c->ready[i]);
    if (c->done[i]) cudaEventDestroy(c->done[i]);
  }
  for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->d_scal) cudaFree(c->d_scal);
  if (c->d_bar) cudaFree(c->d_bar);
  for (int i = 0; i < Comm::kSide; ++i) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->join[i]) cudaEventDestroy(c->join[i]);
  }
  if (c->fork) cudaEventDestroy(c->fork);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  delete c;
  p.comm = nullptr;
  return 0;
}

static int comm_get(Plan& p, Comm** out) {
  if (!p.comm) {
    Comm* c = new Comm();
    p.comm = c;
    SX_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
#ifdef SX_EMU
    emu::mark_eager(c->stream);   // tests/emu/cuda_emu.h, SX_EMU_ADVERSARIAL & 16
#endif
    for (int i = 0; i < 32; ++i) {
      SX_CUDA_CHECK(cudaEventCreate(&c->ready[i]));
      SX_CUDA_CHECK(cudaEventCreate(&c->done[i]));
    }
    SX_CUDA_CHECK(cudaMalloc((void**)&c->d_scal, 64 * sizeof(double)));
    SX_CUDA_CHECK(cudaMalloc((void**)&c->d_bar, 64));
    SX_CUDA_CHECK(cudaMemset(c->d_bar, 0, 64));
    for (int i = 0; i < Comm::kSide; ++i) {
      SX_CUDA_CHECK(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
      SX_CUDA_CHECK(cudaEventCreate(&c->join[i]));
    }
    SX_CUDA_CHECK(cudaEventCreate(&c->fork));
    // copy streams (= copy engines) per exchange: one per peer up to 4 by default (env SX_P2P_SPLIT: 1..8)
    c->nsplit = p.nprocs <= 2 ? 2 : (p.nprocs - 1 < 4 ? p.nprocs - 1 : 4);
    if (const char* e = getenv("SX_P2P_SPLIT")) {
      const int v = atoi(e);
      SX_REQUIRE(v >= 1 && v <= Comm::kSide + 1, "invalid value of the tuning variable SX_P2P_SPLIT (1..8)");
      c->nsplit = v;
    }
    SX_CUDA_CHECK(cudaMallocHost((void**)&c->h_scal, 64 * sizeof(double)));
  }
  *out = p.comm;
  return 0;
}

static int comm_tmark(Comm& c) {   // next timing event on the communication stream
  if (c.ntev == c.tev.size()) {
    cudaEvent_t e;
    SX_CUDA_CHECK(cudaEventCreate(&e));
    c.tev.push_back(e);
  }
  SX_CUDA_CHECK(cudaEventRecord(c.tev[c.ntev++], c.stream));
  return 0;
}
static int comm_resolve(Comm& c) {   // add the elapsed time of every recorded (begin, end) pair
  for (size_t i = 0; i + 1 < c.ntev; i += 2) {
    SX_CUDA_CHECK(cudaEventSynchronize(c.tev[i + 1]));
    float f = 0.f;
    SX_CUDA_CHECK(cudaEventElapsedTime(&f, c.tev[i], c.tev[i + 1]));
    c.ms += f;
  }
  c.ntev = 0;
  return 0;
}

bool comm_ready(const Plan& p) {
  if (p.nprocs == 1) return true;
  if (!p.comm) return false;
#ifndef SX_EMU
  if (p.comm->nccl) return true;
#endif
  return p.comm->a2a != nullptr;
}

// All-to-all-v of contiguous blocks; displacements and counts in complex elements.  `ev` selects the
// event pair (one per buffer in flight).  Asynchronous with respect to the compute stream: the data
// may be read only after exchange_wait(ev).
int exchange_begin(Plan& p, int ev, const cplx* send, cplx* recv, const size_t* sdispl, const size_t* scount,
                   const size_t* rdispl, const size_t* rcount) {
  SX_REQUIRE(p.nprocs > 1, "exchange on a single-rank plan");
  SX_REQUIRE(comm_ready(p), "multi-rank plan without a communicator: call sx_plan_set_comm or sx_plan_set_comm_callbacks");
  SX_REQUIRE(ev >= 0 && ev < 32, "exchange: bad event slot");
  Comm& c = *p.comm;
  if (stage_mark(p, ST_EXCHANGE)) return 1;
  double sent = 0.0;
  for (int r = 0; r < p.nprocs; ++r)
    if (r != p.myrank) sent += (double)scount[r] * sizeof(cplx);
  c.bytes_sent += sent;
  c.exchanges++;
  if (c.a2a) {
    // caller-provided transport: complete the producing kernels, hand over byte displacements
    SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
    std::vector<size_t> sd(p.nprocs), sc(p.nprocs), rd(p.nprocs), rc(p.nprocs);
    for (int r = 0; r < p.nprocs; ++r) {
      sd[r] = sdispl[r] * sizeof(cplx); sc[r] = scount[r] * sizeof(cplx);
      rd[r] = rdispl[r] * sizeof(cplx); rc[r] = rcount[r] * sizeof(cplx);
    }
    const int rc_ = c.a2a(c.user, send, sd.data(), sc.data(), recv, rd.data(), rc.data(), p.nprocs);
    SX_REQUIRE(rc_ == 0, "the all-to-all callback reported an error");
    return 0;
  }
#ifndef SX_EMU
  SX_CUDA_CHECK(cudaEventRecord(c.ready[ev], p.stream));
  SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ready[ev], 0));
  const bool timing = p.timer.on;
  if (timing && comm_tmark(c)) return 1;
  ncclResult_t st = ncclGroupStart();
  for (int q = 0; q < p.nprocs && st == ncclSuccess; ++q) {
    // start with the neighbour so that all pairs are busy (any order is correct inside a group)
    const int r = (p.myrank + q) % p.nprocs, s = (p.myrank - q + p.nprocs) % p.nprocs;
    if (scount[r]) st = ncclSend(send + sdispl[r], scount[r] * 2, ncclDouble, r, c.nccl, c.stream);
    if (st == ncclSuccess && rcount[s]) st = ncclRecv(recv + rdispl[s], rcount[s] * 2, ncclDouble, s, c.nccl, c.stream);
  }
  ncclResult_t st2 = ncclGroupEnd();
  SX_REQUIRE(st == ncclSuccess && st2 == ncclSuccess,
             std::string("NCCL all-to-all failed: ") + ncclGetErrorString(st != ncclSuccess ? st : st2));
  if (timing) {
    if (comm_tmark(c)) return 1;
    if (c.ntev >= 4096 && comm_resolve(c)) return 1;
  }
  SX_CUDA_CHECK(cudaEventRecord(c.done[ev], c.stream));
  p.launches++;
  return 0;
#else
  SX_REQUIRE(false, "the emulated build has no NCCL: install callbacks with sx_plan_set_comm_callbacks");
#endif
}

// The same all-to-all-v with the payload moved by the copy engines: block r of `send` goes straight into rank r's
// receive buffer (peer_dst[r], an IPC-mapped pointer that already includes this rank's displacement there).  The
// copies use no SM and run at NVLink speed next to the transforms of the next field; a one-word NCCL all-reduce
// behind them on the same stream is the completion barrier: when it returns on this rank, every rank has
// finished the copies it enqueued before it, so everything addressed to this rank has landed.  The same chain of
// barriers orders the reuse of a receive buffer in the next substep after its readers of this one.
int exchange_begin_p2p(Plan& p, int ev, const cplx* send, const size_t* sdispl, const size_t* scount, cplx* const* peer_dst) {
#ifndef SX_EMU
  SX_REQUIRE(p.nprocs > 1 && p.comm && p.comm->nccl, "peer-to-peer exchange needs the NCCL communicator (sx_plan_set_comm)");
  SX_REQUIRE(ev >= 0 && ev < 32, "exchange: bad event slot");
  Comm& c = *p.comm;
  if (stage_mark(p, ST_EXCHANGE)) return 1;
  double sent = 0.0;
  for (int r = 0; r < p.nprocs; ++r)
    if (r != p.myrank) sent += (double)scount[r] * sizeof(cplx);
  c.bytes_sent += sent;
  c.exchanges++;
  SX_CUDA_CHECK(cudaEventRecord(c.ready[ev], p.stream));
  SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ready[ev], 0));
  const bool timing = p.timer.on;
  if (timing && comm_tmark(c)) return 1;
  // every block is cut into nsplit pieces that travel on different streams (= different copy engines)
  const int ns = c.nsplit;
  if (ns > 1) {
    SX_CUDA_CHECK(cudaEventRecord(c.fork, c.stream));
    for (int i = 0; i < ns - 1; ++i) SX_CUDA_CHECK(cudaStreamWaitEvent(c.side[i], c.fork, 0));
  }
  for (int q = 1; q <= p.nprocs; ++q) {   // neighbours first, the local block last
    const int r = (p.myrank + q) % p.nprocs;
    if (!scount[r]) continue;
    const size_t piece = (scount[r] + ns - 1) / ns;
    for (int i = 0; i < ns; ++i) {
      const size_t o = (size_t)i * piece;
      if (o >= scount[r]) break;
      const size_t n = scount[r] - o < piece ? scount[r] - o : piece;
      SX_CUDA_CHECK(cudaMemcpyAsync(peer_dst[r] + o, send + sdispl[r] + o, n * sizeof(cplx), cudaMemcpyDeviceToDevice,
                                    i == 0 ? c.stream : c.side[i - 1]));
    }
  }
  for (int i = 0; i < ns - 1; ++i) {
    SX_CUDA_CHECK(cudaEventRecord(c.join[i], c.side[i]));
    SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.join[i], 0));
  }
  ncclResult_t st = ncclAllReduce(c.d_bar, c.d_bar, 1, ncclFloat, ncclSum, c.nccl, c.stream);
  SX_REQUIRE(st == ncclSuccess, std::string("NCCL barrier failed: ") + ncclGetErrorString(st));
  if (timing) {
    if (comm_tmark(c)) return 1;
    if (c.ntev >= 4096 && comm_resolve(c)) return 1;
  }
  SX_CUDA_CHECK(cudaEventRecord(c.done[ev], c.stream));
  p.launches++;
  return 0;
#else
  // emulation: the same stream / event structure as above with one copy stream; the copies go into the peers'
  // shared-memory arenas and the completion barrier is a one-word all-reduce through the caller's callback, run as an
  // operation of the communication stream (tests/emu/cuda_emu.h: in program order by default, deferred with
  // SX_EMU_ADVERSARIAL & 8)
  SX_REQUIRE(p.nprocs > 1 && p.comm && p.comm->allred, "the emulated peer-to-peer exchange needs sx_plan_set_comm_callbacks (barrier)");
  SX_REQUIRE(ev >= 0 && ev < 32, "exchange: bad event slot");
  Comm& c = *p.comm;
  if (stage_mark(p, ST_EXCHANGE)) return 1;
  SX_CUDA_CHECK(cudaEventRecord(c.ready[ev], p.stream));
  SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ready[ev], 0));
  for (int q = 1; q <= p.nprocs; ++q) {
    const int r = (p.myrank + q) % p.nprocs;
    if (!scount[r]) continue;
    if (r != p.myrank) c.bytes_sent += (double)scount[r] * sizeof(cplx);
    SX_CUDA_CHECK(cudaMemcpyAsync(peer_dst[r], send + sdispl[r], scount[r] * sizeof(cplx), cudaMemcpyDeviceToDevice, c.stream));
  }
  c.exchanges++;
  Comm* cp = &c;
  emu::enqueue(c.stream, [cp]() {
    double one = 0.0;
    if (cp->allred(cp->user, &one, 1) != 0) { std::fprintf(stderr, "emulated barrier: the all-reduce callback failed\n"); std::abort(); }
  });
  SX_CUDA_CHECK(cudaEventRecord(c.done[ev], c.stream));
  p.launches++;
  return 0;
#endif
}

// Chunked form of the same transport (the z-chunk pipeline of sx_fused.cu): a round is a list of 2-D block copies
// (rows of one z chunk of every transposed field), optionally closed by the barrier.
int p2p_mark(Plan& p, int slot) {
  SX_REQUIRE(p.comm && slot >= 0 && slot < 32, "p2p_mark: bad slot");
  SX_CUDA_CHECK(cudaEventRecord(p.comm->ready[slot], p.stream));
  return 0;
}

int p2p_round(Plan& p, const int* wait_slots, int nwait, const P2PCopy* cp, int n, bool barrier, int done_slot) {
#ifndef SX_EMU
  SX_REQUIRE(p.nprocs > 1 && p.comm && p.comm->nccl, "peer-to-peer exchange needs the NCCL communicator (sx_plan_set_comm)");
  Comm& c = *p.comm;
  if (stage_mark(p, ST_EXCHANGE)) return 1;
  for (int i = 0; i < nwait; ++i) SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ready[wait_slots[i]], 0));
  const bool timing = p.timer.on;
  if (timing && comm_tmark(c)) return 1;
  const int ns = c.nsplit;
  if (ns > 1 && n > 0) {
    SX_CUDA_CHECK(cudaEventRecord(c.fork, c.stream));
    for (int i = 0; i < ns - 1; ++i) SX_CUDA_CHECK(cudaStreamWaitEvent(c.side[i], c.fork, 0));
  }
  for (int i = 0; i < n; ++i) {   // whole blocks round-robin over the copy streams (= copy engines)
    const P2PCopy& q = cp[i];
    if (q.width == 0 || q.height == 0) continue;
    if (q.remote) c.bytes_sent += (double)q.width * q.height * sizeof(cplx);
    cudaStream_t st = (ns > 1 && i % ns) ? c.side[i % ns - 1] : c.stream;
    SX_CUDA_CHECK(cudaMemcpy2DAsync(q.dst, q.dpitch * sizeof(cplx), q.src, q.spitch * sizeof(cplx), q.width * sizeof(cplx),
                                    q.height, cudaMemcpyDeviceToDevice, st));
  }
  if (ns > 1 && n > 0)
    for (int i = 0; i < ns - 1; ++i) {
      SX_CUDA_CHECK(cudaEventRecord(c.join[i], c.side[i]));
      SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.join[i], 0));
    }
  if (barrier) {
    ncclResult_t st = ncclAllReduce(c.d_bar, c.d_bar, 1, ncclFloat, ncclSum, c.nccl, c.stream);
    SX_REQUIRE(st == ncclSuccess, std::string("NCCL barrier failed: ") + ncclGetErrorString(st));
    c.exchanges++;
  }
  if (timing) {
    if (comm_tmark(c)) return 1;
    if (c.ntev >= 4096 && comm_resolve(c)) return 1;
  }
  if (done_slot >= 0) SX_CUDA_CHECK(cudaEventRecord(c.done[done_slot], c.stream));
  p.launches++;
  return 0;
#else
  SX_REQUIRE(p.nprocs > 1 && p.comm && p.comm->allred, "the emulated peer-to-peer exchange needs sx_plan_set_comm_callbacks (barrier)");
  Comm& c = *p.comm;
  if (stage_mark(p, ST_EXCHANGE)) return 1;
  for (int i = 0; i < nwait; ++i) SX_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ready[wait_slots[i]], 0));
  for (int i = 0; i < n; ++i) {
    const P2PCopy& q = cp[i];
    if (q.width == 0 || q.height == 0) continue;
    if (q.remote) c.bytes_sent += (double)q.width * q.height * sizeof(cplx);
    SX_CUDA_CHECK(cudaMemcpy2DAsync(q.dst, q.dpitch * sizeof(cplx), q.src, q.spitch * sizeof(cplx), q.width * sizeof(cplx),
                                    q.height, cudaMemcpyDeviceToDevice, c.stream));
  }
  if (barrier) {
    Comm* cq = &c;
    emu::enqueue(c.stream, [cq]() {
      double one = 0.0;
      if (cq->allred(cq->user, &one, 1) != 0) { std::fprintf(stderr, "emulated barrier: the all-reduce callback failed\n"); std::abort(); }
    });
    c.exchanges++;
  }
  if (done_slot >= 0) SX_CUDA_CHECK(cudaEventRecord(c.done[done_slot], c.stream));
  p.launches++;
  return 0;
#endif
}

int exchange_wait(Plan& p, int ev) {
  Comm& c = *p.comm;
#ifndef SX_EMU
  if (c.a2a) return 0;
#endif
  SX_CUDA_CHECK(cudaStreamWaitEvent(p.stream, c.done[ev], 0));   // (emulation: a never recorded event is no wait)
  return 0;
}

// Sum of n doubles over all ranks (result on every rank).
int allreduce_sum(Plan& p, double* v, int n) {
  if (p.nprocs == 1) return 0;
  SX_REQUIRE(comm_ready(p), "multi-rank plan without a communicator: call sx_plan_set_comm or sx_plan_set_comm_callbacks");
  SX_REQUIRE(n <= 64, "allreduce_sum: at most 64 values");
  Comm& c = *p.comm;
  if (c.allred) {
    SX_REQUIRE(c.allred(c.user, v, n) == 0, "the all-reduce callback reported an error");
    return 0;
  }
#ifndef SX_EMU
  for (int i = 0; i < n; ++i) c.h_scal[i] = v[i];
  SX_CUDA_CHECK(cudaMemcpyAsync(c.d_scal, c.h_scal, n * sizeof(double), cudaMemcpyHostToDevice, p.stream));
  ncclResult_t st = ncclAllReduce(c.d_scal, c.d_scal, n, ncclDouble, ncclSum, c.nccl, p.stream);
  SX_REQUIRE(st == ncclSuccess, std::string("ncclAllReduce failed: ") + ncclGetErrorString(st));
  SX_CUDA_CHECK(cudaMemcpyAsync(c.h_scal, c.d_scal, n * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  for (int i = 0; i < n; ++i) v[i] = c.h_scal[i];
  return 0;
#else
  SX_REQUIRE(false, "the emulated build has no NCCL: install callbacks with sx_plan_set_comm_callbacks");
#endif
}

}  // namespace sx

using namespace sx;
#define SX_PLAN(pl) \
  if (!(pl)) { sx::set_error("[ERROR] null plan"); return 1; } \
  sx::Plan& p = (pl)->p

extern "C" {

int sx_nccl_unique_id(void* id128) {
#ifndef SX_EMU
  SX_REQUIRE(id128 != nullptr, "sx_nccl_unique_id: null buffer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  ncclResult_t st = ncclGetUniqueId(&id);
  SX_REQUIRE(st == ncclSuccess, std::string("ncclGetUniqueId failed: ") + ncclGetErrorString(st));
  memcpy(id128, &id, sizeof(id));
  return 0;
#else
  (void)id128;
  SX_REQUIRE(false, "the emulated build has no NCCL");
#endif
}

int sx_plan_set_comm(sx_plan* plan, const void* id128) {
  SX_PLAN(plan);
#ifndef SX_EMU
  SX_REQUIRE(id128 != nullptr, "sx_plan_set_comm: null id");
  Comm* c;
  if (comm_get(p, &c)) return 1;
  SX_REQUIRE(c->nccl == nullptr, "sx_plan_set_comm: communicator already set");
  SX_CUDA_CHECK(cudaSetDevice(p.device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t st = ncclCommInitRank(&c->nccl, p.nprocs, id, p.myrank);
  SX_REQUIRE(st == ncclSuccess, std::string("ncclCommInitRank failed: ") + ncclGetErrorString(st));
  return 0;
#else
  (void)id128;
  SX_REQUIRE(false, "the emulated build has no NCCL: use sx_plan_set_comm_callbacks");
#endif
}

int sx_plan_set_comm_callbacks(sx_plan* plan, sx_alltoallv_fn alltoallv, sx_allreduce_fn allreduce, void* user) {
  SX_PLAN(plan);
  SX_REQUIRE(alltoallv != nullptr && allreduce != nullptr, "sx_plan_set_comm_callbacks: null callback");
  Comm* c;
  if (comm_get(p, &c)) return 1;
  c->a2a = alltoallv;
  c->allred = allreduce;
  c->user = user;
  return 0;
}

int sx_plan_p2p_export(sx_plan* plan, int n_inverse, int n_forward, void* handle64) {
  SX_PLAN(plan);
  SX_REQUIRE(handle64 != nullptr && n_inverse > 0 && n_forward > 0, "sx_plan_p2p_export: bad arguments");
  return fused_p2p_export(p, n_inverse, n_forward, handle64);
}

int sx_plan_p2p_import(sx_plan* plan, const void* handles) {
  SX_PLAN(plan);
  SX_REQUIRE(handles != nullptr, "sx_plan_p2p_import: null handles");
  return fused_p2p_import(p, handles);
}

int sx_plan_comm_stats(sx_plan* plan, double* bytes_sent, double* ms, long long* exchanges, int reset) {
  SX_PLAN(plan);
  Comm* c = p.comm;
  if (c && comm_resolve(*c)) return 1;
  if (bytes_sent) *bytes_sent = c ? c->bytes_sent : 0.0;
  if (ms) *ms = c ? c->ms : 0.0;
  if (exchanges) *exchanges = c ? c->exchanges : 0;
  if (c && reset) { c->bytes_sent = 0.0; c->ms = 0.0; c->exchanges = 0; }
  return 0;
}

}  // extern "C"
