// The HD Runge-Kutta substep (include/hd/hd_rkstep{1,2}.f90) on plan-owned, device-resident
// state.  impl=1 composes the substep from the per-operator kernels exactly in the reference's
// order; impl=0 is the fused B200 path (sx_fused.cu).
#include "../../include/specter_b200.h"
#include "sx_plan.h"

namespace sx {

struct HdState {
  cplx* f[10] = {nullptr};  // 0..2 v, 3 pr, 4..6 force, 7..9 RK base (C1..C3)
};

static int hd_state(Plan& p, HdState** out) {
  if (!p.hd) {
    p.hd = new HdState();
    for (int i = 0; i < 10; ++i) {
      SX_CUDA_CHECK(cudaMalloc((void**)&p.hd->f[i], p.csize() * sizeof(cplx)));
      SX_CUDA_CHECK(cudaMemsetAsync(p.hd->f[i], 0, p.csize() * sizeof(cplx), p.stream));
    }
  }
  *out = p.hd;
  return 0;
}

int hd_state_free(Plan& p) {
  if (p.hd) {
    for (int i = 0; i < 10; ++i) if (p.hd->f[i]) cudaFree(p.hd->f[i]);
    delete p.hd;
    p.hd = nullptr;
  }
  return 0;
}

int hd_rkstep2_fused(Plan& p, cplx* const* f, int o, double dt, double nu, const double* zs, const double* ze);

// hd_rkstep2.f90:3-36 composed from stand-alone operators
static int hd_rkstep2_modular(Plan& p, HdState& s, int o, double dt, double nu, const double* zs, const double* ze) {
  cplx *c4, *c5, *c6;
  if (plan_cwork(p, 6, &c4) || plan_cwork(p, 7, &c5) || plan_cwork(p, 8, &c6)) return 1;
  const double rmp = 1.0 / (double)o;
  if (gradre(p, s.f[0], s.f[1], s.f[2], c4, c5, c6)) return 1;
  if (op_fc_filter(p, c4) || op_fc_filter(p, c5) || op_fc_filter(p, c6)) return 1;
  cplx* nl[3] = {c4, c5, c6};
  for (int q = 0; q < 3; ++q) {
    if (op_laplak(p, s.f[q], s.f[q])) return 1;
    if (op_rk_axpy(p, s.f[q], s.f[7 + q], nl[q], s.f[4 + q], dt, nu, rmp)) return 1;
  }
  return v_imposebc_and_project(p, s.f[0], s.f[1], s.f[2], s.f[3], o, zs, ze);
}

}  // namespace sx

using namespace sx;
#define SX_PLAN(pl) \
  if (!(pl)) { sx::set_error("[ERROR] null plan"); return 1; } \
  sx::Plan& p = (pl)->p

extern "C" {

int sx_hd_put_state(sx_plan* plan, const double* vx, const double* vy, const double* vz, const double* pr,
                    const double* fx, const double* fy, const double* fz) {
  SX_PLAN(plan);
  HdState* s;
  if (hd_state(p, &s)) return 1;
  const double* h[7] = {vx, vy, vz, pr, fx, fy, fz};
  const size_t bytes = p.csize() * sizeof(cplx);
  for (int i = 0; i < 7; ++i)
    if (h[i]) SX_CUDA_CHECK(cudaMemcpyAsync(s->f[i], h[i], bytes, cudaMemcpyHostToDevice, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  return 0;
}

int sx_hd_get_state(sx_plan* plan, double* vx, double* vy, double* vz, double* pr) {
  SX_PLAN(plan);
  HdState* s;
  if (hd_state(p, &s)) return 1;
  double* h[4] = {vx, vy, vz, pr};
  const size_t bytes = p.csize() * sizeof(cplx);
  for (int i = 0; i < 4; ++i)
    if (h[i]) SX_CUDA_CHECK(cudaMemcpyAsync(h[i], s->f[i], bytes, cudaMemcpyDeviceToHost, p.stream));
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));
  return 0;
}

int sx_hd_state_ptr(sx_plan* plan, int which, double** dptr) {
  SX_PLAN(plan);
  SX_REQUIRE(which >= 0 && which < 10 && dptr, "sx_hd_state_ptr: which must be 0..9");
  HdState* s;
  if (hd_state(p, &s)) return 1;
  *dptr = reinterpret_cast<double*>(s->f[which]);
  return 0;
}

int sx_hd_rkstep1(sx_plan* plan) {
  SX_PLAN(plan);
  HdState* s;
  if (hd_state(p, &s)) return 1;
  for (int q = 0; q < 3; ++q)
    SX_CUDA_CHECK(cudaMemcpyAsync(s->f[7 + q], s->f[q], p.csize() * sizeof(cplx), cudaMemcpyDeviceToDevice, p.stream));
  return 0;
}

int sx_hd_rkstep2(sx_plan* plan, int o, double dt, double nu, const double v_zsta[2], const double v_zend[2], int impl) {
  SX_PLAN(plan);
  SX_REQUIRE(o >= 1 && o <= p.ord, "sx_hd_rkstep2: substep index o must be in 1..ord");
  SX_REQUIRE(p.Cz > 0, "no-slip walls need a non-periodic z direction (Cz > 0)");
  HdState* s;
  if (hd_state(p, &s)) return 1;
  if (impl == 1) return hd_rkstep2_modular(p, *s, o, dt, nu, v_zsta, v_zend);
  SX_REQUIRE(impl == 0, "sx_hd_rkstep2: impl must be 0 (fused) or 1 (per-operator)");
  return hd_rkstep2_fused(p, s->f, o, dt, nu, v_zsta, v_zend);
}

// One whole time step on HOST arrays (in place).  fx / fy / fz may be NULL: the forcing uploaded by an earlier call
// (or by sx_hd_put_state) is kept -- a constant body force is 3 of the 7 input fields.  The uploads run on a copy
// stream, each followed by the rkstep1 copy of that component; the first kernel of the first substep that reads a
// field waits for that field only (Plan::pre_wait), so the z-inverse of vx runs while vy, vz and pr still travel.
int sx_hd_step_host(sx_plan* plan, double* vx, double* vy, double* vz, double* pr, const double* fx,
                    const double* fy, const double* fz, double dt, double nu, const double v_zsta[2],
                    const double v_zend[2]) {
  SX_PLAN(plan);
  SX_REQUIRE(vx && vy && vz && pr, "sx_hd_step_host: vx, vy, vz, pr must not be NULL");
  HdState* s;
  if (hd_state(p, &s)) return 1;
  if (!p.copy_stream) {
    SX_CUDA_CHECK(cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking));
#ifdef SX_EMU
    emu::mark_eager(p.copy_stream);   // tests/emu/cuda_emu.h, SX_EMU_ADVERSARIAL & 16
#endif
    for (int i = 0; i < 4; ++i) SX_CUDA_CHECK(cudaEventCreateWithFlags(&p.h2d_ev[i], cudaEventDisableTiming));
  }
  const size_t bytes = p.csize() * sizeof(cplx);
  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));   // nothing in flight still reads the state that is overwritten
  const double* fh[3] = {fx, fy, fz};
  for (int i = 0; i < 3; ++i)
    if (fh[i]) SX_CUDA_CHECK(cudaMemcpyAsync(s->f[4 + i], fh[i], bytes, cudaMemcpyHostToDevice, p.copy_stream));
  double* h[4] = {vx, vy, vz, pr};
  for (int i = 0; i < 4; ++i) {
    if (i < 3) {
      SX_CUDA_CHECK(cudaMemcpyAsync(s->f[i], h[i], bytes, cudaMemcpyHostToDevice, p.copy_stream));
      // rkstep1 (hd_rkstep1.f90:4-6)
      SX_CUDA_CHECK(cudaMemcpyAsync(s->f[7 + i], s->f[i], bytes, cudaMemcpyDeviceToDevice, p.copy_stream));
    } else {
      // of p' the step only READS the two wall rows (noslip_z, vboundary.f90:154-211: pr(1,j,i) and pr(nz-Cz,j,i)) before the
      // projection of the first substep overwrites every row (boundary_mod.fpp:371-380): two strided copies of one complex
      // number per pencil instead of the whole field
      const size_t pitch = (size_t)p.nz * sizeof(cplx), npen = (size_t)p.ny * p.nxl;
      // (a kernel reading the rows in place from page-locked memory was tried: 263 K scattered 16-byte PCIe reads took 45 ms
      // against 17 ms for the two strided copy-engine copies below and 19.6 ms for the whole field)
      const int rows[2] = {0, p.nphys() - 1};
      for (int q = 0; q < 2; ++q)
        SX_CUDA_CHECK(cudaMemcpy2DAsync(s->f[3] + rows[q], pitch, reinterpret_cast<const cplx*>(pr) + rows[q], pitch, sizeof(cplx), npen,
                                        cudaMemcpyHostToDevice, p.copy_stream));
    }
    SX_CUDA_CHECK(cudaEventRecord(p.h2d_ev[i], p.copy_stream));
    p.pre_wait[i] = p.h2d_ev[i];
  }
  int rc = 0;
  for (int o = p.ord; o >= 1 && !rc; --o) rc = sx_hd_rkstep2(plan, o, dt, nu, v_zsta, v_zend, 0);
  for (int i = 0; i < 4; ++i) p.pre_wait[i] = nullptr;
  if (rc) {
    cudaStreamSynchronize(p.copy_stream);
    return 1;
  }
  return sx_hd_get_state(plan, vx, vy, vz, pr);
}

}  // extern "C"
