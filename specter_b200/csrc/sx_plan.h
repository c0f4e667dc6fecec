// Internal plan object behind the C ABI (include/specter_b200.h).
#pragma once
#include <string>
#include <vector>
#include "sx_common.cuh"

namespace sx {

struct HdState;  // fused-path device state (sx_rkstep.cu)
struct Fused;    // fused-path work buffers and slab maps (sx_fused.cu)
struct Comm;     // NCCL communicator / transport callbacks (sx_comm.cu)
struct SolverState;  // BOUSS / MHD device state (sx_solvers.cu)

// Per-stage CUDA-event timers (the reference's ffttime/tratime/comtime/conttime counters,
// fftp_mod.fpp:32-37, re-cast per kernel family).
enum Stage {
  ST_OTHER = 0, ST_ZFFT, ST_YFFT, ST_XFFT, ST_EW, ST_REDUCE,          // per-operator path
  ST_ZINV, ST_YINV, ST_XPASS, ST_YFWD, ST_ZFWD_RK, ST_PROJECT, ST_EXCHANGE,  // fused substep
  ST_COUNT
};
const char* stage_name(int id);
struct StageTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ids;
  size_t n = 0;
  double ms[ST_COUNT] = {0};
  long long cnt[ST_COUNT] = {0};
};

struct Plan {
  // configuration (mirrors FCPLAN + grid/kes modules of the reference)
  int nx = 0, ny = 0, nz = 0, Cz = 0, oz = 0, ord = 2;
  double Lx = 1, Ly = 1, Lz = 1;
  int nprocs = 1, myrank = 0, device = 0;
  // slab partition, 1-based inclusive like the reference (fftp.fpp:1154-1184)
  int nxh = 0, ista = 1, iend = 0, ksta = 1, kend = 0, pkend = 0, nxl = 0, nzl = 0;
  double dx = 0, dy = 0, dz = 0, Dkx = 0, Dky = 0, Dkz = 0;
  // host copies
  std::vector<double> h_kx, h_ky, h_kz, h_z, h_dir, h_neu, h_neu2;
  // device tables
  double *d_kx = nullptr, *d_ky = nullptr, *d_kz = nullptr;  // kx is the LOCAL slab kx(ista:iend)
  double* d_kxg = nullptr;                                    // GLOBAL kx(1:nx/2+1)
  double *d_fx = nullptr, *d_fy = nullptr, *d_fz = nullptr;  // fc_filter separable factors
  double *d_z = nullptr, *d_dir = nullptr;
  cplx *tw_x = nullptr, *tw_y = nullptr, *tw_z = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // sx_plan_time_begin/end
  // host-buffer step (sx_hd_step_host): copy stream, one event per uploaded state field, and the events the fused
  // substep still has to wait for before it first reads vx, vy, vz, pr (consumed by sx_fused.cu)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t h2d_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pre_wait[4] = {nullptr, nullptr, nullptr, nullptr};
  // scratch pool (lazily grown): complex spectral-sized and real-sized work arrays
  std::vector<cplx*> cwork;
  std::vector<double*> rwork;
  cplx *xy_T = nullptr, *xy_S = nullptr;   // slab-parallel stand-alone xy transforms: local-plane array and block staging
  double* d_red = nullptr;   // reduction partials
  double* h_red = nullptr;   // pinned host landing zone
  int red_blocks = 0;
  std::string tdir;          // FC-Gram table directory (Neumann tables are loaded on first use)
  // BCPLAN of the three fields at the z walls (setup_bc, boundary_mod.fpp:30-68): v 0 noslip; s 0 constant;
  // b 0 conducting / 1 vacuum
  int v_bczsta = 0, v_bczend = 0, s_bczsta = 0, s_bczend = 0, b_bczsta = 0, b_bczend = 0;
  HdState* hd = nullptr;
  SolverState* bouss = nullptr;
  SolverState* mhd = nullptr;
  SolverState* mhdbouss = nullptr;
  Fused* fused = nullptr;
  Comm* comm = nullptr;
  StageTimer timer;
  int num_sms = 148;                // multiProcessorCount of the device
  int knob_xp = 0, knob_pj = 0;     // 0: bulk-copy x pass / paired projection where they apply; 9: the slot kernels; 10: forced (tests) (env SX_XP, SX_PJ)
  int knob_pf = 13;                 // cp.async prefetch per tile kernel: bit 0 zinv, 1 yinv, 2 yfwd, 3 zfwd (env SX_TILE_PF)
  int knob_tma = 3;                 // bulk-copy (TMA) tile kernels: bit 0 zinv, 1 yinv, 2 yfwd, 3 zfwd (env SX_TMA)
  int knob_tma_min = 256;           // smallest transform length the bulk-copy tile kernels are used for (env SX_TMA_MIN)
  int knob_inv_stages = 2;          // two-stage input ring of the bulk-copy inverse tile kernels: bit 0 zinv, bit 1 yinv (env SX_INV_STAGES)
  int knob_zchunks = 4;             // z chunks of the multi-rank xy stage pipeline (env SX_ZCHUNKS; 1 = unchunked)
  unsigned long long launches = 0;  // kernels launched by this plan (bench "gpu_launches")

  size_t csize() const { return (size_t)nz * ny * nxl; }   // complex elements per spectral field
  size_t rsize() const { return (size_t)nx * ny * nzl; }   // doubles per real field
  int nphys() const { return nz - Cz; }
};

// stage timing: call stage_mark right before a kernel launch
int stage_mark_slow(Plan& p, int id);
inline int stage_mark(Plan& p, int id) { return p.timer.on ? stage_mark_slow(p, id) : 0; }
int stage_flush(Plan& p);

// scratch management
int plan_cwork(Plan& p, int idx, cplx** out);
int plan_rwork(Plan& p, int idx, double** out);

// ---- FFT stage launchers (sx_kernels_fft.cu) ----------------------------------
int launch_zfft(Plan& p, const cplx* in, cplx* out, long npencils, int dir, bool cont,
                double scale_phys, double scale_cont);
int launch_yfft(Plan& p, const cplx* in, cplx* out, int nzc, int nxc, int nz_active, int dir,
                double scale);
int launch_x_c2r(Plan& p, const cplx* spec, double* real, int nzc, int nz_active, double scale);
int launch_x_r2c(Plan& p, const double* real, cplx* spec, int nzc, int nz_active, double scale);
bool fft_size_supported(int n, bool zdir);

// ---- operator launchers (sx_kernels_ops.cu) -------------------------------------
int op_derivk(Plan& p, const cplx* a, cplx* b, int dir);
int op_laplak(Plan& p, const cplx* a, cplx* b);
int op_curlk(Plan& p, const cplx* a, const cplx* b, cplx* c, int dir);
int op_fc_filter(Plan& p, cplx* a);
int op_copy(Plan& p, const cplx* a, cplx* b);
int op_add(Plan& p, cplx* a, const cplx* b);
int op_scale_phys(Plan& p, cplx* a, double s);
int op_scale_copy(Plan& p, const cplx* a, cplx* b, double s);
int fft2d_xy_r2c(Plan& p, const double* r, cplx* out, int nz_active);
int fft2d_xy_c2r(Plan& p, cplx* mixed_destroyed, double* r, int nz_active);
int op_rk_axpy(Plan& p, cplx* v, const cplx* v0, const cplx* nl, const cplx* f, double dt,
               double nu, double rmp);
int op_gradre_products(Plan& p, double* const r[12], double* rx, double* ry, double* rz);
int op_cross_products(Plan& p, const double* a1, const double* a2, const double* a3,
                      const double* b1, const double* b2, const double* b3, double* rx,
                      double* ry, double* rz);
int op_proj_inhomogeneous(Plan& p, cplx* a, cplx* b, cplx* c, cplx* d, cplx* C1, int bctarget);
int op_laplace_z(Plan& p, const cplx* C1, const cplx* C2in, cplx* C2, cplx* C3, int bctarget,
                 int bczsta, int bczend);
int op_pr_combine(Plan& p, cplx* d, const cplx* C1, const cplx* C2, int bctarget);
int op_apply_hom(Plan& p, cplx* a, cplx* b, cplx* c, const cplx* C2, const cplx* C3);
int op_noslip(Plan& p, cplx* vx, cplx* vy, const cplx* pr, int o, double vbx0, double vby0,
              double vbx1, double vby1);
int op_reduce_phys(Plan& p, const cplx* a, const cplx* b, int mode, int row, double scale,
                   double* result);

// ---- composite operators (sx_api.cu) ------------------------------------------------
int fft1d_z_fwd(Plan& p, cplx* a);
int fft1d_z_bwd(Plan& p, const cplx* in, cplx* out, double scale_phys);
int fft3d_r2c(Plan& p, const double* r, cplx* out);
int fft3d_c2r(Plan& p, const cplx* in, double* r);
int gradre(Plan& p, const cplx* a, const cplx* b, const cplx* c, cplx* d, cplx* e, cplx* f);
int prodre(Plan& p, const cplx* a, const cplx* b, const cplx* c, cplx* d, cplx* e, cplx* f);
int sol_project(Plan& p, cplx* a, cplx* b, cplx* c, cplx* d, int bctarget, int bczsta, int bczend);
int v_imposebc_and_project(Plan& p, cplx* vx, cplx* vy, cplx* vz, cplx* pr, int rki, const double* zs,
                           const double* ze);
int energy(Plan& p, const cplx* a, const cplx* b, const cplx* c, int kin, double* out);
int divergence(Plan& p, const cplx* a, const cplx* b, const cplx* c, double* out);
int cross(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* d, const cplx* e, const cplx* f,
          int kin, double* out);
int bouncheck_z(Plan& p, double* bot, double* top, const cplx* a, const cplx* b);
int variance(Plan& p, const cplx* a, int kin, double* out);
int advect(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* d, cplx* e);
int vector(Plan& p, const cplx* a, const cplx* b, const cplx* c, const cplx* d, const cplx* e, const cplx* f,
           cplx* x, cplx* y, cplx* z);
int s_imposebc(Plan& p, cplx* th);
int a_imposebc_and_project(Plan& p, cplx* ax, cplx* ay, cplx* az, cplx* ph);
int hd_state_free(Plan& p);
int solver_states_free(Plan& p);
int op_set_elem(Plan& p, cplx* a, size_t idx, double re, double im);
int load_neumann(Plan& p);
int load_dirichlet(Plan& p, const std::string& tdir);   // dir = A Q^T of (p.Cz, p.oz) into p.h_dir
int fused_free(Plan& p);
int comm_free(Plan& p);
bool comm_ready(const Plan& p);
int exchange_begin(Plan& p, int ev, const cplx* send, cplx* recv, const size_t* sdispl, const size_t* scount,
                   const size_t* rdispl, const size_t* rcount);
int exchange_begin_p2p(Plan& p, int ev, const cplx* send, const size_t* sdispl, const size_t* scount, cplx* const* peer_dst);
int exchange_wait(Plan& p, int ev);
struct P2PCopy {   // 2-D block copy in complex elements
  cplx* dst;
  const cplx* src;
  size_t width, height, spitch, dpitch;
  int remote;
};
int p2p_mark(Plan& p, int slot);
int p2p_round(Plan& p, const int* wait_slots, int nwait, const P2PCopy* cp, int n, bool barrier, int done_slot);
int fused_p2p_export(Plan& p, int nw, int nx, void* handle64);
int fused_p2p_import(Plan& p, const void* handles);
int allreduce_sum(Plan& p, double* v, int n);

}  // namespace sx

// the opaque handle of include/specter_b200.h
struct sx_plan {
  sx::Plan p;
};
