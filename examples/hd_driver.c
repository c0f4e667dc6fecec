/* hd_driver.c -- the HD main loop of the reference's driver written against include/specter_b200.h only.
 *
 * What a maintainer's Fortran driver does through the ISO_C_BINDING module of INTEGRATION.md, in plain C99 (no CUDA, no
 * Python): continue a run from field files (the stat /= 0 branch, specter.fpp:886-912), set the forcing of
 * initialfv.f90:25-31, run the Runge-Kutta loop (specter.fpp:1142-1161 with include/hd/hd_rkstep{1,2}.f90), print the
 * balance.txt / noslip_diagnostic.txt quantities (include/hd/hd_global.f90) every `cstep' steps, write the BIN block
 * (specter.fpp:1005-1053) at the end and append the benchmark.txt row (specter.fpp:1182-1228).
 *
 *   cc -std=c99 -Iinclude examples/hd_driver.c -Lspecter_b200/csrc -lspecter_b200 -Wl,-rpath,$PWD/specter_b200/csrc -o hd_driver
 *   ./hd_driver tdir idir odir nx ny nz Cz oz ord Lx Ly Lz dt nu f0 nsteps cstep ext_in ext_out
 */
#define _POSIX_C_SOURCE 200112L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "specter_b200.h"

#define CHECK(call)                                                    \
  do {                                                                 \
    if ((call) != 0) {                                                 \
      fprintf(stderr, "%s\n  in %s\n", sx_last_error(), #call);        \
      if (plan) sx_plan_destroy(plan);                                 \
      return 1; /* the reference: MPI_FINALIZE; STOP */                \
    }                                                                  \
  } while (0)

static double wall(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char** argv) {
  sx_plan* plan = NULL;
  if (argc != 20) {
    fprintf(stderr, "usage: %s tdir idir odir nx ny nz Cz oz ord Lx Ly Lz dt nu f0 nsteps cstep ext_in ext_out\n", argv[0]);
    return 2;
  }
  const char *tdir = argv[1], *idir = argv[2], *odir = argv[3];
  sx_config cfg;
  cfg.nx = atoi(argv[4]); cfg.ny = atoi(argv[5]); cfg.nz = atoi(argv[6]);
  cfg.Cz = atoi(argv[7]); cfg.oz = atoi(argv[8]); cfg.ord = atoi(argv[9]);
  cfg.Lx = atof(argv[10]); cfg.Ly = atof(argv[11]); cfg.Lz = atof(argv[12]);
  cfg.tdir = tdir; cfg.nprocs = 1; cfg.myrank = 0; cfg.device = -1;
  const double dt = atof(argv[13]), nu = atof(argv[14]), f0 = atof(argv[15]);
  const int nsteps = atoi(argv[16]), cstep = atoi(argv[17]);
  const char *ext_in = argv[18], *ext_out = argv[19];
  const double v_zsta[2] = {0.0, 0.0}, v_zend[2] = {0.0, 0.0};   /* vxzsta, vyzsta / vxzend, vyzend of parameter.inp */

  CHECK(sx_plan_create(&cfg, &plan));
  /* stat /= 0: vx, vy, vz, pr <- idir/<name>.<ext_in>.out (specter.fpp:886-912) */
  CHECK(sx_hd_restart(plan, idir, ext_in, dt));
  /* initialfv.f90:25-31: fx(1,1,1) = f0 nx ny nz on the rank that owns kx = 0 */
  {
    double *fx = NULL, one[2];
    one[0] = f0 * (double)cfg.nx * (double)cfg.ny * (double)cfg.nz; one[1] = 0.0;
    CHECK(sx_hd_state_ptr(plan, 4, &fx));
    CHECK(sx_memcpy_h2d(plan, fx, one, sizeof one));
  }
  double *v[3], *f[3];
  for (int q = 0; q < 3; ++q) { CHECK(sx_hd_state_ptr(plan, q, &v[q])); CHECK(sx_hd_state_ptr(plan, 4 + q, &f[q])); }

  CHECK(sx_plan_stage_timing(plan, 1));
  const double t0 = wall();
  const clock_t c0 = clock();
  for (int t = 1; t <= nsteps; ++t) {
    if (cstep > 0 && (t - 1) % cstep == 0) {   /* hd_global.f90: hdcheck + vdiagnostic, time label (t-1) dt */
      double eng, ens, pot, d[5];
      CHECK(sx_hdcheck(plan, v[0], v[1], v[2], f[0], f[1], f[2], &eng, &ens, &pot));
      CHECK(sx_vdiagnostic(plan, v[0], v[1], v[2], d));
      printf("%.6e %.15e %.15e %.15e | %.6e %.6e %.6e %.6e %.6e\n", (t - 1) * dt, eng, ens, pot, d[0], d[1], d[2], d[3], d[4]);
    }
    CHECK(sx_hd_rkstep1(plan));                                           /* include/hd/hd_rkstep1.f90 */
    for (int o = cfg.ord; o >= 1; --o)                                    /* specter.fpp:1150-1160 */
      CHECK(sx_hd_rkstep2(plan, o, dt, nu, v_zsta, v_zend, 0));           /* include/hd/hd_rkstep2.f90, fused path */
  }
  CHECK(sx_plan_synchronize(plan));
  const double twtime = wall() - t0, tcpu = (double)(clock() - c0) / CLOCKS_PER_SEC;
  CHECK(sx_hd_output(plan, odir, ext_out, dt, 0));                         /* specter.fpp:1005-1053 */
  {
    char path[4096];
    snprintf(path, sizeof path, "%s/benchmark.txt", odir);
    CHECK(sx_benchmark_write(plan, path, nsteps, 1, tcpu, tcpu, twtime));  /* specter.fpp:1182-1228 */
  }
  fprintf(stderr, "hd_driver: %d steps of %dx%dx%d, %llu kernel launches, %.3f s\n", nsteps, cfg.nx, cfg.ny, cfg.nz,
          sx_plan_launch_count(plan), twtime);
  sx_plan_destroy(plan);
  return 0;
}
