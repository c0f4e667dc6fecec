/* specter_b200 -- C ABI of the B200-native backend for SPECTER's per-RK-substep hot path.
 *
 * This header is the drop-in boundary.  Each entry point replaces one Fortran procedure of
 * the reference's fftp / pseudo / boundary modules (cited "ref:" below, paths relative to
 * /root/reference/src).  A thin ISO_C_BINDING module binds these names.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; sx_last_error() gives the
 *    message (the reference prints "[ERROR] ..." and STOPs instead, e.g. vboundary.f90:102).
 *  - one plan per process / GPU (the reference: one MPI rank per GPU, specter.fpp:263-268);
 *    calls on a plan are ordered on the plan's CUDA stream and are not re-entrant.
 *  - array arguments are DEVICE pointers unless the name ends in _host.  Layouts are the
 *    reference's, byte for byte:
 *       spectral / mixed  COMPLEX(GP) a(nz,ny,ista:iend)  -> interleaved (re,im) doubles, z fastest
 *       real              REAL(GP)    r(nx,ny,ksta:kend)  -> doubles, x fastest
 *  - transforms are unnormalised in both directions, as in the reference (fftp.fpp).
 *  - there is no CPU fallback: sx_plan_create fails if no CUDA device is usable.
 */
#ifndef SPECTER_B200_H
#define SPECTER_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sx_plan sx_plan;

typedef struct sx_config {
  int nx, ny, nz;      /* ref: NX,NY,NZ  (Makefile.in:6-8, pseudospec_mod.fpp:18-35) */
  int Cz, oz;          /* continuation points / matching order in z (Makefile.in:16,21) */
  int ord;             /* RK order ORD (Makefile.in:24, module order) */
  double Lx, Ly, Lz;   /* &boxparams (specter.fpp:201-239) */
  const char* tdir;    /* FC-Gram table directory (parameter.inp tdir) */
  int nprocs, myrank;  /* slab decomposition (specter.fpp:288-289); 1,0 for a single GPU */
  int device;          /* CUDA device ordinal; -1 = myrank % device_count (specter.fpp:263-284) */
} sx_config;

const char* sx_last_error(void);
const char* sx_version(void);

/* ref: fcgram_create_plan (fcgram_mod.f90:49-144) + fftp3d_create_plan (fftp.fpp:61-162)
 *      + grid / wavenumber set-up (specter.fpp:683-800) */
int sx_plan_create(const sx_config* cfg, sx_plan** plan);
int sx_plan_destroy(sx_plan* plan);
/* local slab extents, 1-based inclusive as `range` returns them (fftp.fpp:1154-1184) */
int sx_plan_info(const sx_plan* plan, int* ista, int* iend, int* ksta, int* kend, int* pkend);
int sx_range(int n1, int n2, int nprocs, int irank, int* sta, int* end);
/* kernels launched by this plan so far */
unsigned long long sx_plan_launch_count(const sx_plan* plan);
int sx_plan_synchronize(sx_plan* plan);
/* The callee temporaries of the per-operator entries (the reference's automatic arrays C1.., R1.., pseudospec_hd.f90:233-240)
 * are pooled per plan and kept between calls; this frees the pool (it is re-grown on demand), e.g. after a set-up phase
 * and before the fused substep allocates its work fields. */
int sx_plan_release_scratch(sx_plan* plan);
/* CUDA-event timing on the plan's own stream (replaces the GTStart/GTStop wall timers around the
 * RK loop, specter.fpp:985-999): begin records an event, end records another, synchronises and
 * returns the elapsed device time in milliseconds. */
int sx_plan_time_begin(sx_plan* plan);
int sx_plan_time_end(sx_plan* plan, double* ms);
/* Per-kernel-family device timers, the counterpart of the reference's ffttime / tratime / comtime /
 * conttime columns of benchmark.txt (fftp_mod.fpp:32-37, specter.fpp:1182-1228).  While enabled an
 * event is recorded before every kernel launch; sx_plan_stage_times synchronises and returns the
 * accumulated milliseconds and launch counts per stage id (0 .. sx_stage_count()-1). */
int sx_stage_count(void);
const char* sx_stage_name(int id);
int sx_plan_stage_timing(sx_plan* plan, int on);
int sx_plan_stage_times(sx_plan* plan, double* ms, long long* counts, int n);
/* multi-GPU: 128-byte NCCL unique id created on rank 0 and shared by the caller
 * (MPI_BCAST in the Fortran driver, torch.distributed in the Python harness) */
int sx_nccl_unique_id(void* id128);
int sx_plan_set_comm(sx_plan* plan, const void* id128);
/* Peer-to-peer slab exchange for one process per GPU on an NVLink / NVSwitch node.  The receive buffers of the
 * fused substep (n_inverse fields on the way to real space, n_forward on the way back; HD 6/3, BOUSS / ROTBOUSS 8/4,
 * MHD 12/6, MHDBOUSS 14/7) live in one device allocation per rank.  sx_plan_p2p_export creates it and writes its 64-byte CUDA
 * IPC handle; the caller gathers the handles of all ranks (MPI_ALLGATHER in the Fortran driver,
 * torch.distributed in the Python harness) and passes the nprocs x 64 bytes to sx_plan_p2p_import.  From then on
 * every block of the all-to-all-v is written straight into the destination GPU's buffer by the copy engines and
 * NCCL only carries the completion barrier; without these calls the blocks travel as ncclSend/ncclRecv.
 * Replaces the MPI_ISEND/MPI_IRECV ring of fftp/fftp.fpp:478-499, 887-907. */
int sx_plan_p2p_export(sx_plan* plan, int n_inverse, int n_forward, void* handle64);
int sx_plan_p2p_import(sx_plan* plan, const void* handles);
/* Alternative transport (e.g. CUDA-aware MPI_Alltoallv / MPI_Allreduce from the Fortran driver, or
 * torch.distributed in the test-suite).  Buffers are device pointers, displacements and counts are in
 * BYTES per peer rank; the all-reduce sums n doubles in a host array in place.  Return 0 on success. */
typedef int (*sx_alltoallv_fn)(void* user, const void* sendbuf, const size_t* sdispl, const size_t* scount,
                               void* recvbuf, const size_t* rdispl, const size_t* rcount, int nprocs);
typedef int (*sx_allreduce_fn)(void* user, double* inout_host, int n);
int sx_plan_set_comm_callbacks(sx_plan* plan, sx_alltoallv_fn alltoallv, sx_allreduce_fn allreduce, void* user);
/* bytes this rank sent to peers, device milliseconds spent in exchanges (only accumulated while
 * sx_plan_stage_timing is on) and the number of exchanges since the last reset */
int sx_plan_comm_stats(sx_plan* plan, double* bytes_sent, double* ms, long long* exchanges, int reset);

/* device memory helpers so a host language needs no CUDA runtime of its own */
int sx_malloc(sx_plan* plan, size_t bytes, void** dptr);
int sx_free(sx_plan* plan, void* dptr);
int sx_malloc_host(size_t bytes, void** hptr); /* pinned */
int sx_free_host(void* hptr);
int sx_memcpy_h2d(sx_plan* plan, void* dptr, const void* hptr, size_t bytes);
int sx_memcpy_d2h(sx_plan* plan, void* hptr, const void* dptr, size_t bytes);
size_t sx_spectral_bytes(const sx_plan* plan); /* 16*nz*ny*(iend-ista+1) */
size_t sx_real_bytes(const sx_plan* plan);     /* 8*nx*ny*(kend-ksta+1) */

/* ---- fftp module ------------------------------------------------------------ */
/* ref: fftp3d_real_to_complex (fftp.fpp:388-425) */
int sx_fftp3d_real_to_complex(sx_plan* plan, const double* in_real, double* out_spec);
/* ref: fftp3d_complex_to_real (fftp.fpp:789-821); unlike the reference `in` is preserved */
int sx_fftp3d_complex_to_real(sx_plan* plan, const double* in_spec, double* out_real);
/* ref: fftp2d_real_to_complex_xy (fftp.fpp:428-524) */
int sx_fftp2d_real_to_complex_xy(sx_plan* plan, const double* in_real, double* out_mixed);
/* ref: fftp2d_complex_to_real_xy (fftp.fpp:824-919) */
int sx_fftp2d_complex_to_real_xy(sx_plan* plan, const double* in_mixed, double* out_real);
/* ref: fftp1d_real_to_complex_z (fftp.fpp:720-786): FC-Gram continuation + forward z FFT, in place */
int sx_fftp1d_real_to_complex_z(sx_plan* plan, double* inout);
/* ref: fftp1d_complex_to_real_z (fftp.fpp:1060-1094): backward z FFT, in place */
int sx_fftp1d_complex_to_real_z(sx_plan* plan, double* inout);

/* ---- pseudo module ---------------------------------------------------------- */
int sx_derivk(sx_plan* plan, const double* a, double* b, int dir);               /* ref: pseudospec_hd.f90:28-94 */
int sx_laplak(sx_plan* plan, const double* a, double* b);                        /* ref: :97-127 */
int sx_curlk(sx_plan* plan, const double* a, const double* b, double* c, int dir); /* ref: :130-206 */
int sx_fc_filter(sx_plan* plan, double* a);                                      /* ref: :1082-1115 */
/* ref: gradre (pseudospec_hd.f90:209-319): (A.grad)A */
int sx_gradre(sx_plan* plan, const double* a, const double* b, const double* c, double* d, double* e, double* f);
/* ref: prodre (pseudospec_hd.f90:322-402): curl(A) x A */
int sx_prodre(sx_plan* plan, const double* a, const double* b, const double* c, double* d, double* e, double* f);
/* ref: normvec (pseudospec_hd.f90:1238-1286): (a,b,c) *= sqrt(d / energy(a,b,c,kin)); normsca (pseudospec_phd.f90:324-368):
 * a *= sqrt(b / variance(a,kin)); normalize (module_dns.f90:13-42): (fx,fy,fz) *= f0 / sqrt(energy(fx,fy,fz,kin)).
 * What the initial-condition and forcing hooks call after filling a field. */
int sx_normvec(sx_plan* plan, double* a, double* b, double* c, double d, int kin);
int sx_normsca(sx_plan* plan, double* a, double b, int kin);
int sx_normalize(sx_plan* plan, double* fx, double* fy, double* fz, double f0, int kin);
/* diagnostics; the sums and maxima are all-reduced over the plan's communicator INSIDE the library: the result is final
   on every rank (do not add an MPI_REDUCE in the caller; the reference's MPI_REDUCE to rank 0 is subsumed) */
int sx_energy(sx_plan* plan, const double* a, const double* b, const double* c, int kin, double* out);   /* ref: :405-635 */
int sx_divergence(sx_plan* plan, const double* a, const double* b, const double* c, double* out);        /* ref: :1118-1235 */
int sx_cross(sx_plan* plan, const double* a, const double* b, const double* c, const double* d,
             const double* e, const double* f, int kin, double* out);                                     /* ref: :778-940 */
int sx_hdcheck(sx_plan* plan, const double* a, const double* b, const double* c, const double* d,
               const double* e, const double* f, double* eng, double* ens, double* pot);                  /* ref: :943-1005 */

/* ref: advect (pseudospec_phd.f90:23-113): e = A.grad(d) */
int sx_advect(sx_plan* plan, const double* a, const double* b, const double* c, const double* d, double* e);
/* ref: vector (pseudospec_mhd.f90:22-105): (x,y,z) = A x B, A = (a,b,c), B = (d,e,f) */
int sx_vector(sx_plan* plan, const double* a, const double* b, const double* c, const double* d, const double* e,
              const double* f, double* x, double* y, double* z);
/* ref: variance (pseudospec_phd.f90:116-196) */
int sx_variance(sx_plan* plan, const double* a, int kin, double* out);

/* ---- boundary module -------------------------------------------------------- */
/* ref: goto_domain_w_boundaries (boundary_mod.fpp:72-150): backward z transform, physical rows x 1/nz, of one to
 * three fields in place (b, c may be NULL); goto_3d_fourier (:153-194): continuation + forward z transform */
int sx_goto_domain_w_boundaries(sx_plan* plan, double* a, double* b, double* c);
int sx_goto_3d_fourier(sx_plan* plan, double* a, double* b, double* c);
/* ref: sol_project (boundary_mod.fpp:197-402); d returns the potential in the mixed domain */
int sx_sol_project(sx_plan* plan, double* a, double* b, double* c, double* d, int bctarget, int bczsta, int bczend);
/* ref: v_imposebc_and_project (vboundary.f90:67-151); no-slip walls, v_zsta/v_zend = wall (vx,vy) */
int sx_v_imposebc_and_project(sx_plan* plan, double* vx, double* vy, double* vz, double* pr, int rki,
                              const double v_zsta[2], const double v_zend[2]);
/* ref: bouncheck_z (boundary_mod.fpp:681-801); b may be NULL */
int sx_bouncheck_z(sx_plan* plan, double* bot, double* top, const double* a, const double* b);
/* ref: vdiagnostic (vboundary.f90:214-269): out[5] = div, vt0, vtL, vn0, vnL */
int sx_vdiagnostic(sx_plan* plan, const double* a, const double* b, const double* c, double out[5]);

/* ref: s_imposebc (sboundary.f90:67-119) with `constant' walls (s_constant_z :122-165) */
int sx_s_imposebc(sx_plan* plan, double* th);
/* ref: a_imposebc_and_project (bboundary.f90:100-189); wall kinds from sx_setup_bc(plan, "b", ...) (default
 * conducting at both ends; vacuum walls use insulating_z :294-344 with robin_reconstruct): int_conducting_z,
 * sol_project(...,0,0,0), conducting_z with neumann_reconstruct (fcgram_mod.f90:368-511); ph returns the gauge
 * potential in the mixed domain */
int sx_a_imposebc_and_project(sx_plan* plan, double* ax, double* ay, double* az, double* ph);

/* ---- wall BC kinds, stand-alone reconstructions, remaining diagnostics -------- */
/* ref: setup_bc (boundary_mod.fpp:30-68) -> v_setup / s_setup / b_setup.  field = "v" | "s" | "b"; bckind[6] = the
 * strings of parameter.inp for x=0, x=Lx, y=0, y=Ly, z=0, z=Lz: "periodic", "noslip" (v), "constant" (s),
 * "conducting" | "vacuum" (b).  x and y must be periodic.  Defaults without a call: noslip / constant / conducting. */
int sx_setup_bc(sx_plan* plan, const char* field, const char* const bckind[6]);
/* ref: neumann_reconstruct (fcgram_mod.f90:368-511), z branches: f in the mixed domain with the prescribed normal
 * derivative in the wall row; boun 5 (z=0) | 6 (z=Lz); order 1 | 2 */
int sx_neumann_reconstruct(sx_plan* plan, double* f, int boun, int order);
/* ref: robin_reconstruct (fcgram_mod.f90:514-644), z branches: wall value from f' + a f = g; a = device array
 * (ny, ista:iend) of real coefficients, or NULL for khom = sqrt(kx^2+ky^2) (what every caller passes) */
int sx_robin_reconstruct(sx_plan* plan, double* f, int boun, const double* a);
int sx_helicity(sx_plan* plan, const double* a, const double* b, const double* c, double* out); /* ref: pseudospec_hd.f90:638-775 */
int sx_product(sx_plan* plan, const double* a, const double* b, double* out);                   /* ref: pseudospec_phd.f90:199-272 */
/* ref: pscheck (pseudospec_phd.f90:275-321): out = the scalar.txt columns <th^2>, <|k^2 th|^2>, injection */
int sx_pscheck(sx_plan* plan, const double* a, const double* b, double out[3]);
/* ref: maxabs (pseudospec_hd.f90:1008-1079): kin 0 curl, 1 laplacian, 2 the field */
int sx_maxabs(sx_plan* plan, const double* a, const double* b, const double* c, int kin, double* out);
/* ref: mhdcheck (pseudospec_mhd.f90:109-212): out = eng, ens, cur (balance.txt), engk, engm (energy.txt),
 * helk, helm (helicity.txt, hel=1), crh, asq (cross.txt, crs=1) */
int sx_mhdcheck(sx_plan* plan, const double* a, const double* b, const double* c, const double* ma, const double* mb,
                const double* mc, int hel, int crs, double out[9]);
/* ref: robcheck (bboundary.f90:434-602): out = d, e, f, g */
int sx_robcheck(sx_plan* plan, const double* a, const double* b, const double* c, double out[4]);
/* ref: bdiagnostic (bboundary.f90:348-430): the six columns after the time of conducting_diagnostic.txt and / or
 * vacuum_diagnostic.txt, chosen by the wall kinds of sx_setup_bc(plan, "b", ...); *which bit 0 / bit 1 says which */
int sx_bdiagnostic(sx_plan* plan, const double* a, const double* b, const double* c, double conducting[6],
                   double vacuum[6], int* which);
/* ref: sdiagnostic (sboundary.f90:168-210): out = <|th|^2> at z=0 and z=Lz */
int sx_sdiagnostic(sx_plan* plan, const double* a, double out[2]);

/* ---- the RK substep (include/hd/hd_rkstep{1,2}.f90) --------------------------------- */
/* Device-resident HD state owned by the plan.  put/get move whole fields in the reference
 * layout; NULL pointers are skipped. */
int sx_hd_put_state(sx_plan* plan, const double* vx_host, const double* vy_host, const double* vz_host,
                    const double* pr_host, const double* fx_host, const double* fy_host, const double* fz_host);
int sx_hd_get_state(sx_plan* plan, double* vx_host, double* vy_host, double* vz_host, double* pr_host);
/* device pointers of the plan-owned state: which = 0..2 v, 3 pr, 4..6 f, 7..9 RK base C1..C3 */
int sx_hd_state_ptr(sx_plan* plan, int which, double** dptr);
/* ref: hd_rkstep1.f90:4-6 (C1..C3 <- v) */
int sx_hd_rkstep1(sx_plan* plan);
/* ref: hd_rkstep2.f90:3-36, one substep `o` of `ord`.  impl = 0: fused B200 path (default);
 * impl = 1: the same substep composed from the per-operator entry points above. */
int sx_hd_rkstep2(sx_plan* plan, int o, double dt, double nu, const double v_zsta[2], const double v_zend[2], int impl);
/* one full time step (rkstep1 + ord substeps) on host arrays: H2D of v,pr,f, compute, D2H of v,pr.  fx/fy/fz_host may
 * be NULL: the forcing uploaded by an earlier call (or by sx_hd_put_state) stays on the device.  The uploads run on a
 * copy stream and the first substep starts on vx while vy, vz, pr still travel.  Of pr_host only the two wall rows are
 * uploaded: they are all the step reads of it (noslip_z, vboundary.f90:154-211) before every row is overwritten. */
int sx_hd_step_host(sx_plan* plan, double* vx_host, double* vy_host, double* vz_host, double* pr_host,
                    const double* fx_host, const double* fy_host, const double* fz_host, double dt, double nu,
                    const double v_zsta[2], const double v_zend[2]);

/* ---- field files and the output / restart blocks of the driver --------------------------------- */
/* ref: io_write / io_read, mpiio/binary_io.f90:165-223, 89-160: `<dir>/<fname>.<nmb>.out`, raw native reals,
 * Fortran order, global extent (nx, ny, nz-Cz); this rank moves its planes ksta..min(kend, nz-Cz) of the DEVICE
 * real array real_dev(nx, ny, ksta:kend).  io_read zeroes the planes that are not in the file. */
int sx_io_write(sx_plan* plan, const double* real_dev, const char* dir, const char* fname, const char* nmb);
int sx_io_read(sx_plan* plan, double* real_dev, const char* dir, const char* fname, const char* nmb);
/* ref: the BIN block of specter.fpp:1005-1053 on the plan-owned HD state: vx, vy, vz (and wx, wy, wz when
 * outs >= 1) through C = v/N and the 3-D c2r, pr through p = p'/(nx ny dt) and the xy c2r, each to
 * `<odir>/<name>.<ext>.out`.  sx_hd_restart is the stat != 0 branch of specter.fpp:886-912 (files -> state). */
int sx_hd_output(sx_plan* plan, const char* odir, const char* ext, double dt, int outs);
int sx_hd_restart(sx_plan* plan, const char* idir, const char* ext, double dt);
/* The same two blocks for any solver (specter.fpp:1005-1128 and 886-957): solver = "HD" | "BOUSS" | "ROTBOUSS" |
 * "MHD" | "MHDBOUSS" picks the plan-owned state; SCALAR_ adds th, MAGFIELD_ adds ax,ay,az (bx,by,bz for outs >= 1,
 * jx,jy,jz for outs == 2) and ph = ph'/dt */
int sx_output(sx_plan* plan, const char* solver, const char* odir, const char* ext, double dt, int outs);
int sx_restart(sx_plan* plan, const char* solver, const char* idir, const char* ext, double dt);
/* ref: include/<solver>/<solver>_global.f90 on the plan-owned state: hdcheck (hel = 1) or mhdcheck (hel = crs = 1), pscheck,
 * vdiagnostic, bdiagnostic, sdiagnostic, each appending its row to `<odir>/balance.txt', helicity.txt, energy.txt,
 * cross.txt, scalar.txt, noslip_diagnostic.txt, conducting_ / vacuum_diagnostic.txt, scalar_constant_diagnostic.txt in the
 * reference's FORMATs (pseudospec_hd.f90:991-1001, pseudospec_phd.f90:313-318, pseudospec_mhd.f90:189-209,
 * vboundary.f90:260-264, bboundary.f90:400-425, sboundary.f90:201-205); time label (t-1) dt; rank 0 writes */
int sx_global(sx_plan* plan, const char* solver, const char* odir, int t, double dt);
/* ref: the benchmark.txt row of specter.fpp:1182-1228 (non-CUDA column set: nx ny nz nsteps nprocs nth TCPU TOMP
 * TWTIME TFFT TTRA TCOM TCONT TNEU TROB TTOT, seconds per step), appended by rank 0, header when the file is new.
 * TCPU/TOMP/TWTIME are the caller's totals; the T* columns come from the stage timers (sx_plan_stage_timing): the
 * transposition, continuation and wall reconstructions are fused into the transform kernels, so TTRA, TCONT, TNEU,
 * TROB are zero and their time is inside TFFT; TCOM = slab exchanges. */
int sx_benchmark_write(sx_plan* plan, const char* path, int nsteps, int nth, double tcpu, double tomp, double twtime);

/* ---- Boussinesq (include/bouss/bouss_rkstep{1,2}.f90) ------------------------------------ */
/* plan-owned state: which = 0..2 v, 3 pr, 4..6 f, 7..9 C1..C3, 10 th, 11 fs, 12 C7 */
int sx_bouss_put_state(sx_plan* plan, const double* vx_host, const double* vy_host, const double* vz_host,
                       const double* pr_host, const double* th_host, const double* fx_host, const double* fy_host,
                       const double* fz_host, const double* fs_host);
int sx_bouss_get_state(sx_plan* plan, double* vx_host, double* vy_host, double* vz_host, double* pr_host,
                       double* th_host);
int sx_bouss_state_ptr(sx_plan* plan, int which, double** dptr);
/* ref: bouss_rkstep1.f90:4-7 */
int sx_bouss_rkstep1(sx_plan* plan);
/* ref: bouss_rkstep2.f90:3-59 (gradre + advect, buoyancy / heat-current coupling, RK update, no-slip projection,
 * s_imposebc, fc_filter and the theta 3-D round trip).  impl as in sx_hd_rkstep2. */
int sx_bouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                     const double v_zsta[2], const double v_zend[2], int impl);

/* ---- vector-potential MHD (include/mhd/mhd_rkstep{1,2}.f90) ------------------------------- */
/* ref: include/rotbouss/rotbouss_rkstep2.f90:3-56 on the BOUSS state (rotbouss_rkstep1 = bouss_rkstep1):
 * omega[3] = the rotation vector of the `rotation' namelist; no theta filter / round trip at the end */
int sx_rotbouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double kappa, double xmom, double xtemp,
                        const double omega[3], const double v_zsta[2], const double v_zend[2], int impl);

/* plan-owned state: which = 0..2 v, 3 pr, 4..6 f, 7..9 C1..C3, 10..12 a, 13 ph, 14..16 m, 17..19 C9..C11 */
int sx_mhd_put_state(sx_plan* plan, const double* vx_host, const double* vy_host, const double* vz_host,
                     const double* pr_host, const double* ax_host, const double* ay_host, const double* az_host,
                     const double* fx_host, const double* fy_host, const double* fz_host, const double* mx_host,
                     const double* my_host, const double* mz_host);
int sx_mhd_get_state(sx_plan* plan, double* vx_host, double* vy_host, double* vz_host, double* pr_host,
                     double* ax_host, double* ay_host, double* az_host, double* ph_host);
int sx_mhd_state_ptr(sx_plan* plan, int which, double** dptr);
/* ref: mhd_rkstep1.f90:4-9 */
int sx_mhd_rkstep1(sx_plan* plan);
/* ref: mhd_rkstep2.f90:3-84 (B = curl A + b0, J = curl B, prodre - Lorentz force, EMF, RK update, no-slip
 * projection, conducting-wall gauge projection).  b0 may be NULL (no uniform field). */
int sx_mhd_rkstep2(sx_plan* plan, int o, double dt, double nu, double mu, const double b0[3], int impl);

/* ---- MHDBOUSS (include/mhdbouss/mhdbouss_rkstep{1,2}.f90) ------------------------------------------
 * state slots: the MHD ones (0..2 v, 3 pr, 4..6 f, 7..9 C1..C3, 10..12 a, 13 ph, 14..16 m, 17..19 C9..C11),
 * 20 th, 21 fs, 22 C7.  impl 0 = fused slab-parallel path (the MHD passes + a scalar-advection x pass), 1 = per-operator. */
int sx_mhdbouss_put_state(sx_plan* plan, const double* vx_host, const double* vy_host, const double* vz_host,
                          const double* pr_host, const double* ax_host, const double* ay_host, const double* az_host,
                          const double* th_host, const double* fx_host, const double* fy_host, const double* fz_host,
                          const double* mx_host, const double* my_host, const double* mz_host, const double* fs_host);
int sx_mhdbouss_get_state(sx_plan* plan, double* vx_host, double* vy_host, double* vz_host, double* pr_host,
                          double* ax_host, double* ay_host, double* az_host, double* ph_host, double* th_host);
int sx_mhdbouss_state_ptr(sx_plan* plan, int which, double** dptr);
int sx_mhdbouss_rkstep1(sx_plan* plan);
int sx_mhdbouss_rkstep2(sx_plan* plan, int o, double dt, double nu, double mu, double kappa, double xmom,
                        double xtemp, const double b0[3], int impl);

/* ---- BOOTS regridder (tools/boots.fpp) ---------------------------------------------
 * Prolongation of field files of an old grid (nxt, nyt, nzt physical rows, the `regrid' namelist) to the grid
 * (nx, ny, nzp = nz-Cz physical rows) by zero padding in Fourier space; the non-periodic z direction is continued with
 * the FC-Gram table A<Czt>-<ozt>.dat / Q<ozt>.dat of tdir on the old grid.  One GPU (`device', -1 = 0); x and y sizes
 * are powers of two, the z sizes are free. */
/* ref: boots.fpp:176-182: continuation points of the old grid (the table that is needed) and of the new grid */
int sx_boots_points(int nzt, int nzp, int* Czt, int* Czn);
/* ref: boots.fpp:249-303 for one field: in_host (nxt, nyt, nzt) -> out_host (nx, ny, nzp), Fortran order */
int sx_boots_regrid(int device, int nxt, int nyt, int nzt, int ozt, const char* tdir, int nx, int ny, int nzp,
                    const double* in_host, double* out_host);
/* ref: the file loop of boots.fpp:228-312: fnlist = names separated by ';', read as `idir/name' (bmangle = 0),
 * written as `odir/name_P<nx>-<ny>-<nzp>' (i5.5 each) */
int sx_boots_files(int device, const char* idir, const char* odir, const char* tdir, const char* fnlist, int nxt, int nyt,
                   int nzt, int ozt, int nx, int ny, int nzp);
/* kernels launched by the regridder so far */
unsigned long long sx_boots_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
