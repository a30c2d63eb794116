/*
 * flutas_b200.h -- C ABI of libflutas_b200.so: the B200-native (sm_100a) pressure-Poisson path of FluTAS.
 *
 * Every entry point is what a thin Fortran `iso_c_binding` shim with the reference's subroutine names
 * and argument lists forwards to (the shim is shown in INTEGRATION.md).  Conventions:
 *   - all reals are FP64, all arrays Fortran column-major, passed as the address of their FIRST element
 *     (c_loc(arr)), with the halo widths the reference uses:
 *         p(0:n1+1, 0:n2+1, 0:n3+1)                      u,v,w(1-nh_u:n1+nh_u, 1-nh_u:n2+nh_u, 1-nh_u:n3+nh_u)
 *         dzci, dzfi(1-nh_d:n3+nh_d)                     lambdaxy(n_z(1), n_z(2))      a,b,c(n_z(3))
 *   - field pointers (p,u,v,w) may be DEVICE pointers (cudaMalloc / flutas_b200_alloc / managed memory;
 *     the call is then asynchronous on the library stream, like a kernel launch) or plain HOST pointers
 *     (the library stages them through device buffers and returns after the result is back on the host).
 *     Small coefficient arrays (lambdaxy, a, b, c, dzci, dzfi, rhsb*) may live on either side.
 *   - return value 0 = success; anything else is an error whose text is flutas_b200_last_error().
 *     The reference has no status arguments on this path (errors print and stop, src/fft.f90:879-883);
 *     the Fortran shim does the same with a non-zero return.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
 */
#ifndef FLUTAS_B200_H
#define FLUTAS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUTAS_B200_OK 0
#define FLUTAS_B200_ERR_CUDA 1        /* a CUDA runtime call failed / no device                     */
#define FLUTAS_B200_ERR_UNSUPPORTED 2 /* BC pair / transform length not available on this path      */
#define FLUTAS_B200_ERR_ARG 3         /* inconsistent arguments                                     */

/* Library version string, e.g. "flutas_b200 0.1 (sm_100a)". */
const char *flutas_b200_version(void);
const char *flutas_b200_last_error(void);

/* Replaces the GPU binding of initmpi (src/initmpi.f90:59-63: cudaSetDevice(local rank)) and records
 * the slab decomposition: `rank` of `nranks` owns z-planes [rank*ng3/nranks, (rank+1)*ng3/nranks)
 * (= the reference's _DECOMP_X layout with dims_in = (1, nranks), src/initmpi.f90:87-104). */
int flutas_b200_init(int device, int rank, int nranks);

/* CUDA stream (cudaStream_t) all subsequent work is enqueued on; NULL = the legacy default stream. */
int flutas_b200_set_stream(void *cuda_stream);

/* Device memory for fields (the reference uses CUDA managed arrays, main__single_phase.f90:157-163;
 * a gfortran host maps these with c_f_pointer). */
void *flutas_b200_alloc(size_t bytes);              /* device memory (cudaMalloc)                                        */
void *flutas_b200_alloc_managed(size_t bytes);      /* managed memory: host code may touch it too (cudaMallocManaged)    */
void flutas_b200_free(void *ptr);
int flutas_b200_memcpy(void *dst, const void *src, size_t bytes);   /* any direction, synchronous */
int flutas_b200_synchronize(void);

/* fftini, src/fft.f90:24-157.  n_x, n_y: x- and y-pencil sizes (mod_common_mpi); bcxy(0:1,2) as four
 * characters x0,x1,y0,y1; c_or_f(2).  Fills arrplan(2,2) (Fortran order: fwd-x, bwd-x, fwd-y, bwd-y)
 * with opaque handles and normfft exactly as :71,87,125,150.  Supported: 'c' with every BC pair of the
 * reference's table (src/fft.f90:233-291: PP -> R2HC/HC2R, NN -> REDFT10/01, DD -> RODFT10/01,
 * ND -> REDFT11, DN -> RODFT11; the reference's own GPU path stops at PP/NN/DD, :879-883) and even
 * lengths whose half factors into 2,3,5. */
int flutas_b200_fftini(const int n_x[3], const int n_y[3], const char bcxy[4], const char c_or_f[2],
                       void *arrplan[4], double *normfft);

/* fftend, src/fft.f90:159-179. */
int flutas_b200_fftend(void *arrplan[4]);

/* fft(plan,arr), src/fft.f90:181-193 (dfftw_execute_r2r(plan,arr,arr), :188-190): one batched, unnormalised,
 * in-place r2r transform of a dense pencil array, input and output in FFTW's element order (half-complex
 * r0..r_{n/2}, i_{n/2-1}..i_1 for R2HC; k = 0..n-1 for the DCT/DST kinds).  plan = one of the four arrplan
 * handles (its direction and axis are the handle's) or a stand-alone plan; n = extents of arr (the x pencil
 * for x plans, the y pencil for y plans).  flutas_b200_solver does NOT go through this call -- its stages
 * hand the spectrum over in the kernels' own slot order -- it serves hosts that keep the reference's
 * solver_cpu.f90 (its transposes and Thomas loops) and the per-kind parity tests. */
int flutas_b200_fft(void *plan, const int n[3], double *arr);

/* Stand-alone plan with the arguments fftini hands to fftw_plan_guru_r2r (src/fft.f90:75-86,113-124,
 * interface src/fftw.f90:15-36): rank-1 transform (n, is), two howmany dimensions (hm_n, hm_is), in
 * place (os = is).  Served layouts: x (is = 1, hm_is = (n, n*hm_n[0])) and y (is = hm_n[0], hm_is =
 * (1, hm_n[0]*n)).  kind = FFTW's integer code (src/fftw.f90:41-61): R2HC 0, HC2R 1, REDFT01 4, REDFT10 5,
 * REDFT11 6, RODFT01 8, RODFT10 9, RODFT11 10.  flutas_b200_plan_dims returns the array extents the plan
 * was made for; flutas_b200_destroy_plan is dfftw_destroy_plan (src/fft.f90:165-175).
 * libflutas_b200_fftw.so (csrc/fftw_seam.cpp) exports the reference's own FFTW symbols on top of these
 * three calls, so that fft.f90 / fftw.f90 link unchanged without FFTW. */
int flutas_b200_plan_r2r(int n, int is, const int hm_n[2], const int hm_is[2], int kind, void **plan);
int flutas_b200_plan_dims(void *plan, int n[3]);
int flutas_b200_destroy_plan(void *plan);

/* solver_cpu / solver_gpu, src/solver_cpu.f90:20-115, src/solver_gpu.f90:31-472.
 * n = local x-pencil interior size; lambdaxy, a, b, c as returned by initsolver (CPU, i.e. FFTW
 * half-complex eigenvalue order, src/initsolver.f90:87-93,136-139); bcz(0:1); c_or_f(3) (must be 'c').
 * Solves in place on the interior of p; halos are left untouched.  lambdaxy/a/b/c are cached on the
 * device at the first call (they are constant after initsolver); call flutas_b200_solver_invalidate
 * if they ever change. */
int flutas_b200_solver(const int n[3], void *const arrplan[4], double normfft, const double *lambdaxy,
                       const double *a, const double *b, const double *c, const char bcz[2],
                       const char c_or_f[3], double *p);
int flutas_b200_solver_invalidate(void *const arrplan[4]);

/* ---- multi-GPU: z-slab decomposition over the GPUs of one NVSwitch box ------------------------------
 * Layout = the reference's _DECOMP_X with dims_in = (1, nranks): every rank holds p(0:ng1+1, 0:ng2+1,
 * 0:ng3/nranks+1).  x and y transforms are rank-local; the two exchanges of the reference GPU slab
 * path (transpose_xc_to_z / transpose_z_to_xc, src/2decomp/transpose_x_to_z.f90:15-166,
 * transpose_z_to_x.f90:15-163, called at src/solver_gpu.f90:150-153,179-182) are either
 *   (a) a host-supplied all-to-all on device buffers (NCCL / CUDA-aware MPI) -- the pack is fused into
 *       the y-transform store and the unpack into the inverse y-transform load, or
 *   (b) direct NVLink stores into the peers' buffers (CUDA IPC mappings) from inside the y-transform
 *       and Thomas kernels, ordered by two flag barriers per solve: no collective call on the data path.
 * The all-to-all callback moves `bytes_per_peer` bytes from sendbuf + q*bytes_per_peer to rank q and
 * receives rank q's block at recvbuf + q*bytes_per_peer, enqueued on `cuda_stream`; returns 0 on success. */
typedef int (*flutas_b200_alltoall_fn)(void *ctx, const void *sendbuf, void *recvbuf, size_t bytes_per_peer,
                                       void *cuda_stream);
int flutas_b200_set_alltoall(flutas_b200_alltoall_fn fn, void *ctx);

/* (b): export allocates this rank's exchange memory for the plan and writes an opaque blob of
 * flutas_b200_p2p_handle_bytes() bytes; the host all-gathers the blobs (MPI_Allgather / torch.distributed)
 * and passes all nranks of them, in rank order, to attach.  p2p_errors returns the number of barrier
 * time-outs observed (0 when healthy). */
size_t flutas_b200_p2p_handle_bytes(void);
int flutas_b200_p2p_export(void *const arrplan[4], const int n_local[3], void *blob);
int flutas_b200_p2p_attach(void *const arrplan[4], const void *blobs);
int flutas_b200_p2p_errors(void *const arrplan[4]);

/* Schedule of the slab solver (optional; every rank must pass the same values).  pipe_chunks: the forward half runs
 * pipelined over that many k-chunks -- the x transform of chunk c+1 on part of the SMs while the y transform + NVLink
 * stores of chunk c use the rest (0 or 1 = off); pipe_xsm_pct: share of the SMs given to the x kernels (10..90);
 * zcopy: 1 = backward exchange through the copy engines, 0 = fused into the z kernel's stores, -1 = built-in rule.
 * A negative pipe_* value leaves that knob unchanged.  Defaults: FLUTAS_B200_PIPE / _PIPE_XSM / _ZCOPY or off/50/-1. */
int flutas_b200_slab_config(int pipe_chunks, int pipe_xsm_pct, int zcopy);

/* 1 (default): with the direct NVLink exchange attached and an exactly uniform z grid, the z stage of the slab solver
 * runs as a DISTRIBUTED tridiagonal solve -- two rank-local sweeps around a 2P x 2P interface system per column -- and
 * the two all-to-all transposes of the reference's slab path (src/solver_gpu.f90:150-153,179-182) disappear: 32 bytes
 * per column cross NVLink instead of 14 bytes per point.  0: always transpose.  Same value on every rank. */
int flutas_b200_slab_distributed_z(int on);
int flutas_b200_slab_last_distributed(void *const arrplan[4]);   /* 1: the last solver_slab call ran the distributed z solve */

/* solver on a z-slab: as flutas_b200_solver, with n = the LOCAL interior size (ng1, ng2, ng3/nranks) and
 * lambdaxy_global = lambdaxy(ng1, ng2) for the whole x-y plane (the shim all-gathers the (ng1, ng2/nranks)
 * windows initsolver produces on each rank, src/initsolver.f90:87-93; done once).  Collective. */
int flutas_b200_solver_slab(const int n[3], void *const arrplan[4], double normfft, const double *lambdaxy_global,
                            const double *a, const double *b, const double *c, const char bcz[2],
                            const char c_or_f[3], double *p);

/* fillps, src/fillps.f90:16-69 (with _CONSTANT_COEFFS_POISSON: the result is multiplied by rho0). */
int flutas_b200_fillps(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                       const double *dzfi, double dti, double rho0, const double *u, const double *v,
                       const double *w, double *p);

/* updt_rhs_b, src/bound.f90:829-944 (cell-centred).  cbc(0:1,3) as six characters; rhsbx(ny,nz,0:1), rhsby(nx,nz,0:1),
 * rhsbz(nx,ny,0:1) with the LOCAL sizes.  On a z-slab decomposition (flutas_b200_init with nranks > 1) the x and y
 * faces are applied on every rank and the z faces only where the reference's neighbour is MPI_PROC_NULL (:915,929):
 * the bottom one on rank 0, the top one on rank nranks-1. */
int flutas_b200_updt_rhs_b(int nx, int ny, int nz, const char cbc[6], const double *rhsbx,
                           const double *rhsby, const double *rhsbz, double *p);

/* correc, src/correc.f90:16-81 (constant-coefficient branch).  `rho` is accepted for signature
 * compatibility and never dereferenced (it is a (0,0,0)-sized dummy in single-phase runs). */
int flutas_b200_correc(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                       const double *dzci, double dt, double rho0, const double *p, double *u, double *v,
                       double *w, const double *rho);

/* pres_sp_src, src/source.f90:311-346 (single phase): predictor pressure-gradient term from the OLD pressure,
 * u += f_t12*( -(pold(i+1)-pold(i))*dxi )*rho0i (and v, w); called at main__single_phase.f90 before the solve. */
int flutas_b200_pres_sp_src(int nx, int ny, int nz, double f_t12, double dxi, double dyi, double dzi, int nh_d,
                            int nh_u, const double *dzci, double rho0i, const double *pold, double *u, double *v,
                            double *w);

/* pres_tw_src, src/source.f90:247-309 (two phase), _CONSTANT_COEFFS_POISSON branch (:288-293): split pressure
 * gradient with the extrapolated pressure (1 + f_t12/f_t12_o) p - (f_t12/f_t12_o) pold.  rho(0:,0:,0:) as p. */
int flutas_b200_pres_tw_src(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                            const double *dzci, double rho0i, double f_t12, double f_t12_o, const double *p,
                            const double *pold, const double *rho, double *u, double *v, double *w);

/* Pressure bookkeeping loops of the RK sub-step (interior of the halo-1 arrays, halos untouched):
 * mode 0: pold = p (src/apps/single_phase/main__single_phase.f90:693-699); mode 1: p = pold + p (:734-740). */
int flutas_b200_pold_update(int nx, int ny, int nz, int mode, double *p, double *pold);

/* chkdiv, src/chkdiv.f90:18-69.  Returns this rank's divtot / divmax (the caller all-reduces across
 * ranks exactly where the reference calls MPI_ALLREDUCE, :64-65).  Synchronous. */
int flutas_b200_chkdiv(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                       const double *dzfi, const double *u, const double *v, const double *w,
                       double *divtot, double *divmax);

/* load(io,filename,n,fld), src/load.f90:21-89: restart files (fldp.bin, fldu.bin, ...) are headerless raw FP64 in GLOBAL
 * column-major (ng1,ng2,ng3) order whatever the decomposition.  io = 'r' | 'w'; n[3] = this rank's block, start[3] its
 * 0-based global offset (2DECOMP xstart-1); `fld` = host or device pointer to an array with `nh` halo cells per side
 * (nh = 0: the reference's dense fld(n1,n2,n3)).  'r' fails if the file is missing or its size is not 8*ng1*ng2*ng3
 * (the reference aborts, :40-66); 'w' sizes the file and writes the block (any rank order, no truncation race). */
int flutas_b200_load(char io, const char *filename, const int ng[3], const int n[3], const int start[3], int nh,
                     double *fld);

/* boundp, src/bound.f90:146-225 (with set_bc :227-420 and updthalo :946-1110), halo width nh_p = 1:
 * ghost cells of a cell-centred scalar (p, pold) for the reference's _DECOMP_X layout.  Same step order as
 * the reference (y halo, z halo, x faces, y faces, z faces), so edges and corners are bit-identical.
 * cbc(0:1,3) as six characters, bc(0:1,3) the six boundary values, dl(3), dzc(1-nh_d:) / dzf(1-nh_d:)
 * (dzf is unused, like in the reference).  The reference's `halo` argument (MPI datatypes) has no
 * counterpart.  On a z-slab decomposition (flutas_b200_init with nranks > 1) the z halo planes come from
 * the neighbouring ranks through the callback registered with flutas_b200_set_halo_exchange; ranks at a
 * non-periodic wall apply the boundary condition instead (neighbour = MPI_PROC_NULL in the reference). */
int flutas_b200_boundp(const char cbc[6], const int n[3], const double bc[6], int nh_d, int nh_p,
                       const double dl[3], const double *dzc, const double *dzf, double *p);

/* z-halo exchange for flutas_b200_boundp on several ranks: send `count` doubles from send_lo to rank `lo`
 * and from send_hi to rank `hi`, receive the same amounts into recv_lo (from lo) and recv_hi (from hi),
 * all device pointers, ordered on `stream`; lo / hi = -1 means no neighbour (skip that pair).
 * Must return 0 on success.  Counterpart of the MPI_SENDRECV pair of updthalo (src/bound.f90:1098-1103). */
typedef int (*flutas_b200_halo_fn)(void *ctx, const double *send_lo, const double *send_hi, double *recv_lo,
                                   double *recv_hi, size_t count, int lo, int hi, void *stream);
int flutas_b200_set_halo_exchange(flutas_b200_halo_fn fn, void *ctx);

/* bounduvw, src/bound.f90:17-144 (+ set_bc :227-646, updthalo :946-1110, outflow :649-773): ghost cells of the staggered
 * velocity for any halo width nh_u (1 = 'cen', 3 = 'fll'; up to 8).  cbc(0:1,3,3) as 18 characters and bc(0:1,3,3) as 18
 * values in Fortran storage order (side, direction, component); isoutflow(0:1,3) as six ints (0 / non-zero);
 * dl(3); dzc, dzf(1-nh_d:).  The reference's `halo` argument (MPI datatypes) has no counterpart.  Same step order as the
 * reference, so edges and corners are bit-identical.  On a z-slab decomposition the nh_u z-halo planes per side travel
 * through the flutas_b200_set_halo_exchange callback; the z walls and z outflow are applied on ranks 0 / nranks-1 only. */
int flutas_b200_bounduvw(const char cbc[18], const int n[3], const double bc[18], int nh_d, int nh_u, const int isoutflow[6],
                         const double dl[3], const double *dzc, const double *dzf, double *u, double *v, double *w);

/* chkdt_sp / chkdt_tw, src/chkdt.f90:24-199: the field reduction only (:62-85 = :150-173) -- this rank's maximum of the
 * convective inverse time scales dtix, dtiy, dtiz.  The caller all-reduces it (MPI_MAX, :92,183) and evaluates the scalar
 * formulas with its physical parameters (:93-110, :184-196) as the reference does.  Synchronous. */
int flutas_b200_chkdt(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u, const double *dzci,
                      const double *dzfi, const double *u, const double *v, const double *w, double *dti);

/* Optional per-stage device timing (CUDA events on the library stream), the counterpart of the
 * reference's named profiler clocks (src/profiler.f90:103-205; labels "SOLVER", "CORREC", ...).
 * Stage ids 0..count-1 have names ("xfft_fwd", "yfft_fwd", "thomas_z", "yfft_bwd", "xfft_bwd", "fillps",
 * "correc", "exchange_fwd", "exchange_bwd").  profile_read synchronises, fills ms_sum[count] and
 * counts[count] with the totals since the previous read, and clears them. */
int flutas_b200_profile_enable(int on);
int flutas_b200_profile_stage_count(void);
const char *flutas_b200_profile_stage_name(int id);
int flutas_b200_profile_read(double *ms_sum, long *counts);

/* Number of kernels launched by this library since load (bench.py's gpu_launches claim). */
long flutas_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLUTAS_B200_H */
