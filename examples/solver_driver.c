/* solver_driver.c -- the main loop of the reference's driver for any of its solvers (SOLVER = HD | BOUSS | ROTBOUSS |
 * MHD | MHDBOUSS, src/Makefile.in:27), written against include/specter_b200.h only (plain C99: no CUDA, no Python).
 *
 * The sequence of specter.fpp with the solver's include texts: wall kinds of the vector potential from the `magbound'
 * namelist (setup_bc, boundary_mod.fpp:30-68), continue a run from field files (stat /= 0, specter.fpp:886-957), the
 * uniform body force of initialfv.f90:25-31, then per time step the global quantities every `cstep' steps
 * (include/<solver>/<solver>_global.f90 -> balance.txt, helicity.txt, energy.txt, cross.txt, scalar.txt and the wall
 * diagnostics, in the reference's FORMATs) and the Runge-Kutta loop (specter.fpp:1142-1161 with
 * include/<solver>/<solver>_rkstep{1,2}.f90), and the BIN output block at the end (specter.fpp:1005-1128).
 *
 *   cc -std=c99 -Iinclude examples/solver_driver.c -Lspecter_b200/csrc -lspecter_b200 -Wl,-rpath,$PWD/specter_b200/csrc -o solver_driver
 *   ./solver_driver SOLVER tdir idir odir nx ny nz Cz oz ord Lx Ly Lz dt nu kappa mu f0 nsteps cstep ext_in ext_out bz0kind bzLkind
 *                   bx0 by0 bz0 omegax omegay omegaz
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "specter_b200.h"

#define CHECK(call)                                                    \
  do {                                                                 \
    if ((call) != 0) {                                                 \
      fprintf(stderr, "%s\n  in %s\n", sx_last_error(), #call);        \
      if (plan) sx_plan_destroy(plan);                                 \
      return 1; /* the reference: MPI_FINALIZE; STOP */                \
    }                                                                  \
  } while (0)

int main(int argc, char** argv) {
  sx_plan* plan = NULL;
  if (argc != 31) {
    fprintf(stderr, "usage: %s SOLVER tdir idir odir nx ny nz Cz oz ord Lx Ly Lz dt nu kappa mu f0 nsteps cstep ext_in ext_out "
                    "bz0kind bzLkind bx0 by0 bz0 omegax omegay omegaz\n", argv[0]);
    return 2;
  }
  const char *solver = argv[1], *tdir = argv[2], *idir = argv[3], *odir = argv[4];
  sx_config cfg;
  cfg.nx = atoi(argv[5]); cfg.ny = atoi(argv[6]); cfg.nz = atoi(argv[7]);
  cfg.Cz = atoi(argv[8]); cfg.oz = atoi(argv[9]); cfg.ord = atoi(argv[10]);
  cfg.Lx = atof(argv[11]); cfg.Ly = atof(argv[12]); cfg.Lz = atof(argv[13]);
  cfg.tdir = tdir; cfg.nprocs = 1; cfg.myrank = 0; cfg.device = -1;
  const double dt = atof(argv[14]), nu = atof(argv[15]), kappa = atof(argv[16]), mu = atof(argv[17]), f0 = atof(argv[18]);
  const int nsteps = atoi(argv[19]), cstep = atoi(argv[20]);
  const char *ext_in = argv[21], *ext_out = argv[22];
  const char* const bkind[6] = {"periodic", "periodic", "periodic", "periodic", argv[23], argv[24]};   /* magbound namelist */
  const double b0[3] = {atof(argv[25]), atof(argv[26]), atof(argv[27])};                               /* uniformb namelist */
  const double omega[3] = {atof(argv[28]), atof(argv[29]), atof(argv[30])};                            /* rotation namelist */
  const double v_zsta[2] = {0.0, 0.0}, v_zend[2] = {0.0, 0.0}, xmom = 1.0, xtemp = 1.0;

  const int is_hd = !strcmp(solver, "HD"), is_bouss = !strcmp(solver, "BOUSS"), is_rot = !strcmp(solver, "ROTBOUSS"),
            is_mhd = !strcmp(solver, "MHD"), is_mhdb = !strcmp(solver, "MHDBOUSS");
  if (!(is_hd || is_bouss || is_rot || is_mhd || is_mhdb)) {
    fprintf(stderr, "unknown SOLVER %s\n", solver);
    return 2;
  }

  CHECK(sx_plan_create(&cfg, &plan));
  if (is_mhd || is_mhdb) CHECK(sx_setup_bc(plan, "b", bkind));
  CHECK(sx_restart(plan, solver, idir, ext_in, dt));
  {  /* initialfv.f90:25-31: fx(1,1,1) = f0 nx ny nz; slot 4 of every solver's state is fx */
    double *fx = NULL, one[2];
    one[0] = f0 * (double)cfg.nx * (double)cfg.ny * (double)cfg.nz; one[1] = 0.0;
    if (is_hd) CHECK(sx_hd_state_ptr(plan, 4, &fx));
    else if (is_bouss || is_rot) CHECK(sx_bouss_state_ptr(plan, 4, &fx));
    else if (is_mhd) CHECK(sx_mhd_state_ptr(plan, 4, &fx));
    else CHECK(sx_mhdbouss_state_ptr(plan, 4, &fx));
    CHECK(sx_memcpy_h2d(plan, fx, one, sizeof one));
  }

  for (int t = 1; t <= nsteps; ++t) {
    if (cstep > 0 && (t - 1) % cstep == 0) CHECK(sx_global(plan, solver, odir, t, dt));    /* <solver>_global.f90 */
    if (is_hd) CHECK(sx_hd_rkstep1(plan));                                                  /* <solver>_rkstep1.f90 */
    else if (is_bouss || is_rot) CHECK(sx_bouss_rkstep1(plan));
    else if (is_mhd) CHECK(sx_mhd_rkstep1(plan));
    else CHECK(sx_mhdbouss_rkstep1(plan));
    for (int o = cfg.ord; o >= 1; --o) {                                                    /* <solver>_rkstep2.f90 */
      if (is_hd) CHECK(sx_hd_rkstep2(plan, o, dt, nu, v_zsta, v_zend, 0));
      else if (is_bouss) CHECK(sx_bouss_rkstep2(plan, o, dt, nu, kappa, xmom, xtemp, v_zsta, v_zend, 0));
      else if (is_rot) CHECK(sx_rotbouss_rkstep2(plan, o, dt, nu, kappa, xmom, xtemp, omega, v_zsta, v_zend, 0));
      else if (is_mhd) CHECK(sx_mhd_rkstep2(plan, o, dt, nu, mu, b0, 0));
      else CHECK(sx_mhdbouss_rkstep2(plan, o, dt, nu, mu, kappa, xmom, xtemp, b0, 0));
    }
  }
  CHECK(sx_plan_synchronize(plan));
  CHECK(sx_output(plan, solver, odir, ext_out, dt, 0));                                     /* specter.fpp:1005-1128 */
  fprintf(stderr, "solver_driver: %s, %d steps of %dx%dx%d, %llu kernel launches\n", solver, nsteps, cfg.nx, cfg.ny, cfg.nz,
          sx_plan_launch_count(plan));
  sx_plan_destroy(plan);
  return 0;
}
