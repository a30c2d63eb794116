// kernels.cuh -- sm_100a kernels of the pressure-Poisson path.
//
//   xfft_kernel   x-line transforms (lines contiguous; p with halo <-> dense work array)   fft.f90:75-86 / solver_cpu.f90:59,89,93
//   yfft_kernel   y-line transforms (stride n1; staged as (TB x N2) smem tiles, in place)  fft.f90:113-124 / solver_cpu.f90:65,86
//   thomas_*      z tridiagonal solves, one thread per (i,j) column                        solver_cpu.f90:117-223
//   fillps/correc/chkdiv stencils                                                          fillps.f90:42-61, correc.f90:49-73, chkdiv.f90:46-58
//
// All FP64, all bandwidth-bound; no tensor cores by design (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>

#include "geom.cuh"
#include "tile_fft.cuh"

namespace fb {

template <int TB, bool ROT, bool FWD>
__device__ __forceinline__ void tile_transform(double* tile, const LinePlan& P, const cpx* wM, int lane, int worker,
                                               int nworkers) {
  const bool iv = kind_is_iv(P.kind);                     // ND / DN (REDFT11 / RODFT11): pre- and post-twiddle instead of a split
  const TileAcc<TB, ROT> acc{tile, P.M, lane};
  if (FWD) {
    if (iv) { iv_pre<true>(P.M, P.kind, P.wQ, P.pos, worker, nworkers, acc); __syncthreads(); }
    for (int q = 0; q < P.npass; ++q) {
      fft_pass<TB, ROT, true>(tile, P, wM, q, lane, worker, nworkers);
      __syncthreads();
    }
    if (iv) iv_post<true>(P.M, P.kind, P.wN, P.pos, worker, nworkers, acc);
    else split_fwd<TB, ROT>(tile, P, lane, worker, nworkers);
  } else {
    if (iv) iv_pre<false>(P.M, P.kind, P.wQ, P.pos, worker, nworkers, acc);
    else merge_bwd<TB, ROT>(tile, P, lane, worker, nworkers);
    for (int q = P.npass - 1; q >= 0; --q) {
      __syncthreads();
      fft_pass<TB, ROT, false>(tile, P, wM, q, lane, worker, nworkers);
    }
    if (iv) { __syncthreads(); iv_post<false>(P.M, P.kind, P.wN, P.pos, worker, nworkers, acc); }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// x direction.  One block = TB consecutive lines.  FWD: physical -> spectral rows, BWD: the reverse
// (times `scale`, = normfft on the last stage, solver_cpu.f90:93).
// Load/store: a warp handles 32 consecutive elements e of every line; the slot of e is computed once
// and reused for the TB lines, whose TB loads are all in flight together.
template <int TB, bool FWD>
__global__ void __launch_bounds__(256, 3) xfft_kernel(LinePlan P, const double* __restrict__ src, LineGeom gs,
                                                   double* __restrict__ dst, LineGeom gd, double scale) {
  extern __shared__ double tile[];
  const int N = P.N, M = P.M, kind = P.kind;
  cpx* s_w = reinterpret_cast<cpx*>(tile + (size_t)N * TB);
  long* s_off = reinterpret_cast<long*>(s_w + M);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long line0 = (long)blockIdx.x * TB;
  const int nlive = (int)min((long)TB, gs.nlines - line0);

  if (tid < TB) s_off[tid] = line_offset(gs, min(line0 + tid, gs.nlines - 1));
  stage_twiddles(s_w, P.wM, M, tid, nthr);
  __syncthreads();

  for (int e = tid; e < N; e += nthr) {
    int m, part;
    double sgn = 1.0;
    if (FWD) elem_to_slot(kind, N, e, m, part, sgn);
    else { part = (e >= M); m = e - part * M; }
    double v[TB];
#pragma unroll
    for (int L = 0; L < TB; ++L) v[L] = __ldcs(src + s_off[L] + e);
#pragma unroll
    for (int L = 0; L < TB; ++L) tile[taddr<TB, true>(m, part, M, L)] = (L < nlive) ? sgn * v[L] : 0.0;
  }
  __syncthreads();

  const int lane = tid & (TB - 1), worker = tid / TB, nworkers = nthr / TB;
  tile_transform<TB, true, FWD>(tile, P, s_w, lane, worker, nworkers);

  if (tid < TB) s_off[tid] = line_offset(gd, min(line0 + tid, gd.nlines - 1));
  __syncthreads();
  for (int e = tid; e < N; e += nthr) {
    int m, part;
    double sgn = 1.0;
    if (!FWD) elem_to_slot(kind, N, e, m, part, sgn);
    else { part = (e >= M); m = e - part * M; }
    const double f = sgn * scale;
    double v[TB];
#pragma unroll
    for (int L = 0; L < TB; ++L) v[L] = tile[taddr<TB, true>(m, part, M, L)];
#pragma unroll
    for (int L = 0; L < TB; ++L)
      if (L < nlive) __stcs(dst + s_off[L] + e, f * v[L]);
  }
}

// ------------------------------------------------------------------------------------------------
// y direction, in place on the dense work array W(n1, N, n3).  One block = lanes i0..i0+TB-1 of one
// k plane, all N rows.  Global accesses are TB*8 contiguous bytes per row, YU rows in flight per thread.
template <int TB, bool FWD>
__global__ void __launch_bounds__(256, 3) yfft_kernel(LinePlan P, double* __restrict__ W, int n1, int ntile_i,
                                                      SpecGeom sg) {
  constexpr int YU = 8;
  extern __shared__ double tile[];
  const int N = P.N, M = P.M, kind = P.kind;
  cpx* s_w = reinterpret_cast<cpx*>(tile + (size_t)N * TB);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ti = blockIdx.x % ntile_i;
  const long k = blockIdx.x / ntile_i;
  const int i0 = ti * TB;
  const int lane = tid & (TB - 1), worker = tid / TB, nworkers = nthr / TB;
  const bool live = (i0 + lane) < n1;
  const int il = i0 + (live ? lane : 0);
  double* pbase = W + (long)n1 * N * k + il;            // physical side: element e at pbase[e*n1]
  double* sbase = spec_base(sg, il, N, k);              // spectral side: row r at sbase[r*sg.n1l]
  const double* base = FWD ? pbase : sbase;
  const long ldi = FWD ? (long)n1 : (long)sg.n1l;

  stage_twiddles(s_w, P.wM, M, tid, nthr);
  for (int e0 = worker; e0 < N; e0 += YU * nworkers) {
    double v[YU];
#pragma unroll
    for (int u = 0; u < YU; ++u) {
      const int e = e0 + u * nworkers;
      v[u] = (e < N) ? __ldcs(base + (long)e * ldi) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < YU; ++u) {
      const int e = e0 + u * nworkers;
      if (e < N) {
        int m, part;
        double sgn = 1.0;
        if (FWD) elem_to_slot(kind, N, e, m, part, sgn);
        else { part = (e >= M); m = e - part * M; }
        tile[taddr<TB, false>(m, part, M, lane)] = live ? sgn * v[u] : 0.0;
      }
    }
  }
  __syncthreads();

  tile_transform<TB, false, FWD>(tile, P, s_w, lane, worker, nworkers);

  if (!live) return;
  double* obase = FWD ? sbase : pbase;
  const long ldo = FWD ? (long)sg.n1l : (long)n1;
  for (int e0 = worker; e0 < N; e0 += YU * nworkers) {
    double v[YU];
#pragma unroll
    for (int u = 0; u < YU; ++u) {
      const int e = e0 + u * nworkers;
      int m, part;
      double sgn = 1.0;
      if (!FWD) elem_to_slot(kind, N, min(e, N - 1), m, part, sgn);
      else { const int ee = min(e, N - 1); part = (ee >= M); m = ee - part * M; }
      v[u] = sgn * tile[taddr<TB, false>(m, part, M, lane)];
    }
#pragma unroll
    for (int u = 0; u < YU; ++u) {
      const int e = e0 + u * nworkers;
      if (e < N) __stcs(obase + (long)e * ldo, v[u]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// z direction, generic version: one thread per (i,j) column of W(ncol, nz), coalesced over columns.
// Follows dgtsv_homebrewed / gaussel_periodic (solver_cpu.f90:147-223) with the forward-sweep
// intermediates spilled to scratch fields D (and P2 for periodic z).
// `pin`: the (kx,ky)=(0,0) column of an all-Neumann/periodic problem is singular; the reference
// leaves its additive constant to round-off (z ~ 1e-16 pivot, :209-214 and :175-176).  We define the
// gauge instead: p(nz) = 0 for that column (the reference's own `z == 0` branch).
__global__ void thomas_generic_kernel(long ncol, int nz, const double* __restrict__ a, const double* __restrict__ b,
                                      const double* __restrict__ c, const double* __restrict__ lam,
                                      double* __restrict__ W, double* __restrict__ D, int singular) {
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const double l = lam[col];
  const bool pin = singular && (l == 0.0);
  const int n = nz;
  double z = 1.0 / (b[0] + l);
  double d = c[0] * z;
  double p = W[col] * z;
  D[col] = d;
  W[col] = p;
  for (int k = 1; k < n - 1; ++k) {
    const double ak = a[k];
    z = 1.0 / ((b[k] + l) - ak * d);
    d = c[k] * z;
    p = (W[col + k * ncol] - ak * p) * z;
    D[col + k * ncol] = d;
    W[col + k * ncol] = p;
  }
  z = (b[n - 1] + l) - a[n - 1] * d;
  {
    const double num = W[col + (long)(n - 1) * ncol] - a[n - 1] * p;
    p = (pin || z == 0.0) ? 0.0 : num / z;
    W[col + (long)(n - 1) * ncol] = p;
  }
  for (int k = n - 2; k >= 0; --k) {
    p = W[col + k * ncol] - D[col + k * ncol] * p;
    W[col + k * ncol] = p;
  }
}

__global__ void thomas_periodic_generic_kernel(long ncol, int nz, const double* __restrict__ a,
                                               const double* __restrict__ b, const double* __restrict__ c,
                                               const double* __restrict__ lam, double* __restrict__ W,
                                               double* __restrict__ D, double* __restrict__ P2, int singular) {
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const double l = lam[col];
  const bool pin = singular && (l == 0.0);
  const int n = nz, m = nz - 1;                 // reduced system size (solver_cpu.f90:168-174)
  double z = 1.0 / (b[0] + l);
  double d = c[0] * z;
  double p1 = W[col] * z;
  double p2 = -a[0] * z;
  D[col] = d; W[col] = p1; P2[col] = p2;
  for (int k = 1; k < m - 1; ++k) {
    const double ak = a[k];
    z = 1.0 / ((b[k] + l) - ak * d);
    d = c[k] * z;
    p1 = (W[col + k * ncol] - ak * p1) * z;
    p2 = (0.0 - ak * p2) * z;
    D[col + k * ncol] = d; W[col + k * ncol] = p1; P2[col + k * ncol] = p2;
  }
  {
    const int k = m - 1;
    z = (b[k] + l) - a[k] * d;
    const double n1 = W[col + (long)k * ncol] - a[k] * p1, n2 = -c[k] - a[k] * p2;
    p1 = (z != 0.0) ? n1 / z : 0.0;
    p2 = (z != 0.0) ? n2 / z : 0.0;
    W[col + (long)k * ncol] = p1; P2[col + (long)k * ncol] = p2;
  }
  const double p1m = p1, p2m = p2;
  for (int k = m - 2; k >= 0; --k) {
    const double dk = D[col + k * ncol];
    p1 = W[col + k * ncol] - dk * p1;
    p2 = P2[col + k * ncol] - dk * p2;
    W[col + k * ncol] = p1; P2[col + k * ncol] = p2;
  }
  const double num = W[col + (long)m * ncol] - c[m] * p1 - a[m] * p1m;
  const double den = (b[m] + l) + c[m] * p2 + a[m] * p2m;
  const double pn = pin ? 0.0 : num / den;
  W[col + (long)m * ncol] = pn;
  for (int k = 0; k < m; ++k) W[col + k * ncol] = W[col + k * ncol] + P2[col + k * ncol] * pn;
  (void)n;
}

// lambda in the internal spectral layout: lam_int(rx, ry) = lambdaxy(mode_x[rx], mode_y[ry])
__global__ void permute_lambda_kernel(int n1, int n2, const double* __restrict__ lam, const int* __restrict__ mx,
                                      const int* __restrict__ my, double* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n1 * n2) return;
  const int ry = (int)(idx / n1), rx = (int)(idx - (long)ry * n1);
  out[idx] = lam[mx[rx] + (long)n1 * my[ry]];
}

// One axis of a dense (n1, n2, n3) array between the kernels' spectral slot order and FFTW's order (map[slot] = FFTW index):
// to_fftw: dst(.., map[s], ..) = src(.., s, ..); else dst(.., s, ..) = src(.., map[s], ..).  flutas_b200_fft only.
__global__ void fft_permute_kernel(int n1, int n2, long n3, int axis, int to_fftw, const int* __restrict__ map,
                                   const double* __restrict__ src, double* __restrict__ dst) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n1 * n2 * n3) return;
  const long jk = idx / n1;
  const int i = (int)(idx - jk * n1);
  const long k = jk / n2;
  const int j = (int)(jk - k * n2);
  const long other = axis == 0 ? idx - i + map[i] : ((long)k * n2 + map[j]) * n1 + i;
  if (to_fftw) dst[other] = src[idx]; else dst[idx] = src[other];
}

// ------------------------------------------------------------------------------------------------
// stencils.  u,v,w have halo nh_u (lower bound 1-nh_u), p has halo 1 (lower bound 0).
struct StencilGeom {
  int nx, ny, nz, nh_u;
  long su1, su2, sp1, sp2;   // leading dimensions: u(su1, su2, *), p(sp1, sp2, *)
};
__device__ __forceinline__ long uidx(const StencilGeom& g, int i, int j, int k) {
  return (long)(i + g.nh_u - 1) + g.su1 * ((long)(j + g.nh_u - 1) + g.su2 * (long)(k + g.nh_u - 1));
}
__device__ __forceinline__ long pidx(const StencilGeom& g, int i, int j, int k) {
  return (long)i + g.sp1 * ((long)j + g.sp2 * (long)k);
}

// fillps.f90:50-57 (+ updt_rhs_b, bound.f90:858-942, folded in when rhsb* are given)
__global__ void __launch_bounds__(256) fillps_kernel(StencilGeom g, double dtidxi, double dtidyi, double dti,
                                                     const double* __restrict__ dzfi, double rho0,
                                                     const double* __restrict__ u, const double* __restrict__ v,
                                                     const double* __restrict__ w, double* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k);
  const double uc = u[c], vc = v[c], wc = w[c];
  // evaluation order exactly as written in fillps.f90:50-57, no FMA contraction (bit-exact with the oracle)
  const double tz = __dmul_rn(__dmul_rn(__dsub_rn(wc, w[c - g.su1 * g.su2]), dti), dzfi[k]);
  const double ty = __dmul_rn(__dsub_rn(vc, v[c - g.su1]), dtidyi);
  const double tx = __dmul_rn(__dsub_rn(uc, u[c - 1]), dtidxi);
  const double val = __dadd_rn(__dadd_rn(tz, ty), tx);
  p[pidx(g, i, j, k)] = __dmul_rn(val, rho0);
}

// fillps.f90:50-57, two points per thread with aligned 16-byte accesses (see correc_vec2_kernel for the alignment argument):
// pair (i, i+1), i even; u(i-1) of the first point is the one extra scalar load.  Bit-exact with fillps_kernel.
__global__ void __launch_bounds__(256) fillps_vec2_kernel(StencilGeom g, double dtidxi, double dtidyi, double dti,
                                                          const double* __restrict__ dzfi, double rho0,
                                                          const double* __restrict__ u, const double* __restrict__ v,
                                                          const double* __restrict__ w, double* __restrict__ p) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);          // even, 0 .. nx
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k), q = pidx(g, i, j, k);
  const bool lo = (i >= 1), hi = (i + 1 <= g.nx);
  const double2 uc = *reinterpret_cast<const double2*>(u + c);
  const double2 vc = *reinterpret_cast<const double2*>(v + c), vm = *reinterpret_cast<const double2*>(v + c - g.su1);
  const double2 wc = *reinterpret_cast<const double2*>(w + c), wm = *reinterpret_cast<const double2*>(w + c - g.su1 * g.su2);
  const double um = lo ? u[c - 1] : 0.0;
  const double dz = dzfi[k];
  double2 out;
  {
    const double tz = __dmul_rn(__dmul_rn(__dsub_rn(wc.x, wm.x), dti), dz);
    const double ty = __dmul_rn(__dsub_rn(vc.x, vm.x), dtidyi);
    const double tx = __dmul_rn(__dsub_rn(uc.x, um), dtidxi);
    out.x = __dmul_rn(__dadd_rn(__dadd_rn(tz, ty), tx), rho0);
  }
  {
    const double tz = __dmul_rn(__dmul_rn(__dsub_rn(wc.y, wm.y), dti), dz);
    const double ty = __dmul_rn(__dsub_rn(vc.y, vm.y), dtidyi);
    const double tx = __dmul_rn(__dsub_rn(uc.y, uc.x), dtidxi);
    out.y = __dmul_rn(__dadd_rn(__dadd_rn(tz, ty), tx), rho0);
  }
  if (lo && hi) *reinterpret_cast<double2*>(p + q) = out;
  else if (hi) p[q + 1] = out.y;
  else if (lo) p[q] = out.x;
}

// bound.f90:858-942 for one rank owning all six faces; which faces apply is encoded in rhsb pointers
__global__ void updt_rhs_b_kernel(StencilGeom g, const double* __restrict__ rx, double* __restrict__ p) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  // the three face families run as three launches in x, y, z order, like the reference's three loops
  if (idx < (long)ny * nz) {
    const int j = (int)(idx % ny) + 1, k = (int)(idx / ny) + 1;
    p[pidx(g, 1, j, k)] += rx[idx];
    p[pidx(g, nx, j, k)] += rx[idx + (long)ny * nz];
  }
}
__global__ void updt_rhs_b_y_kernel(StencilGeom g, const double* __restrict__ ry, double* __restrict__ p) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  if (idx < (long)nx * nz) {
    const int i = (int)(idx % nx) + 1, k = (int)(idx / nx) + 1;
    p[pidx(g, i, 1, k)] += ry[idx];
    p[pidx(g, i, ny, k)] += ry[idx + (long)nx * nz];
  }
}
// sides: bit 0 = this rank owns the bottom wall (bottom == MPI_PROC_NULL, bound.f90:915), bit 1 = the top wall (:929)
__global__ void updt_rhs_b_z_kernel(StencilGeom g, const double* __restrict__ rz, double* __restrict__ p, int sides) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = g.nx, ny = g.ny, nz = g.nz;
  if (idx < (long)nx * ny) {
    const int i = (int)(idx % nx) + 1, j = (int)(idx / nx) + 1;
    if (sides & 1) p[pidx(g, i, j, 1)] += rz[idx];
    if (sides & 2) p[pidx(g, i, j, nz)] += rz[idx + (long)nx * ny];
  }
}

// correc.f90:57-60 (constant-coefficient branch; rho is never touched)
__global__ void __launch_bounds__(256) correc_kernel(StencilGeom g, double factori, double factorj, double dt,
                                                     const double* __restrict__ dzci, double rho0i,
                                                     const double* __restrict__ p, double* __restrict__ u,
                                                     double* __restrict__ v, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k), q = pidx(g, i, j, k);
  const double pc = p[q];
  // correc.f90:57-60, evaluated left to right without FMA contraction
  u[c] = __dsub_rn(u[c], __dmul_rn(__dmul_rn(factori, __dsub_rn(p[q + 1], pc)), rho0i));
  v[c] = __dsub_rn(v[c], __dmul_rn(__dmul_rn(factorj, __dsub_rn(p[q + g.sp1], pc)), rho0i));
  w[c] = __dsub_rn(w[c], __dmul_rn(__dmul_rn(__dmul_rn(dt, dzci[k]), __dsub_rn(p[q + g.sp1 * g.sp2], pc)), rho0i));
}

// correc.f90:57-60, two points per thread: every row of u,v,w,p starts 16-byte aligned when nx is even (sanity.f90:155)
// and the base pointers are, so the pair (i, i+1) with i even -- 0-based storage element i+nh_u-1 resp. i -- is one
// aligned 16-byte access in all four arrays.  The first pair of a row holds the halo i = 0, the last one i = nx+1: those
// elements are loaded but never stored.  Same arithmetic per element as correc_kernel (bit-exact).
__global__ void __launch_bounds__(256) correc_vec2_kernel(StencilGeom g, double factori, double factorj, double dt,
                                                          const double* __restrict__ dzci, double rho0i,
                                                          const double* __restrict__ p, double* __restrict__ u,
                                                          double* __restrict__ v, double* __restrict__ w) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);          // even, 0 .. nx
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k), q = pidx(g, i, j, k);
  const double2 pc = *reinterpret_cast<const double2*>(p + q);
  const double2 py = *reinterpret_cast<const double2*>(p + q + g.sp1);
  const double2 pz = *reinterpret_cast<const double2*>(p + q + g.sp1 * g.sp2);
  const double px2 = (i + 2 <= g.nx + 1) ? p[q + 2] : 0.0;
  double2 uu = *reinterpret_cast<double2*>(u + c), vv = *reinterpret_cast<double2*>(v + c), ww = *reinterpret_cast<double2*>(w + c);
  const double fk = __dmul_rn(dt, dzci[k]);
  const bool lo = (i >= 1), hi = (i + 1 <= g.nx);                    // which of the two points are interior
  if (lo) {
    uu.x = __dsub_rn(uu.x, __dmul_rn(__dmul_rn(factori, __dsub_rn(pc.y, pc.x)), rho0i));
    vv.x = __dsub_rn(vv.x, __dmul_rn(__dmul_rn(factorj, __dsub_rn(py.x, pc.x)), rho0i));
    ww.x = __dsub_rn(ww.x, __dmul_rn(__dmul_rn(fk, __dsub_rn(pz.x, pc.x)), rho0i));
  }
  if (hi) {
    uu.y = __dsub_rn(uu.y, __dmul_rn(__dmul_rn(factori, __dsub_rn(px2, pc.y)), rho0i));
    vv.y = __dsub_rn(vv.y, __dmul_rn(__dmul_rn(factorj, __dsub_rn(py.y, pc.y)), rho0i));
    ww.y = __dsub_rn(ww.y, __dmul_rn(__dmul_rn(fk, __dsub_rn(pz.y, pc.y)), rho0i));
  }
  if (lo && hi) {
    *reinterpret_cast<double2*>(u + c) = uu; *reinterpret_cast<double2*>(v + c) = vv; *reinterpret_cast<double2*>(w + c) = ww;
  } else if (hi) { u[c + 1] = uu.y; v[c + 1] = vv.y; w[c + 1] = ww.y; }
  else if (lo) { u[c] = uu.x; v[c] = vv.x; w[c] = ww.x; }
}

// source.f90:311-346 (pres_sp_src): u += f_t12*( -(pold(ip)-pold(i))*dxi )*rho0i, left to right, no FMA contraction
__global__ void __launch_bounds__(256) pres_sp_src_kernel(StencilGeom g, double f_t12, double dxi, double dyi,
                                                          const double* __restrict__ dzci, double rho0i,
                                                          const double* __restrict__ pold, double* __restrict__ u,
                                                          double* __restrict__ v, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k), q = pidx(g, i, j, k);
  const double pc = pold[q];
  u[c] = __dadd_rn(u[c], __dmul_rn(__dmul_rn(f_t12, -__dmul_rn(__dsub_rn(pold[q + 1], pc), dxi)), rho0i));
  v[c] = __dadd_rn(v[c], __dmul_rn(__dmul_rn(f_t12, -__dmul_rn(__dsub_rn(pold[q + g.sp1], pc), dyi)), rho0i));
  w[c] = __dadd_rn(w[c], __dmul_rn(__dmul_rn(f_t12, -__dmul_rn(__dsub_rn(pold[q + g.sp1 * g.sp2], pc), dzci[k])), rho0i));
}

// source.f90:247-309 (pres_tw_src), _CONSTANT_COEFFS_POISSON branch :288-293
__global__ void __launch_bounds__(256) pres_tw_src_kernel(StencilGeom g, double dxi, double dyi, const double* __restrict__ dzci,
                                                          double rho0i, double f_t12, double f1, double f2,
                                                          const double* __restrict__ p, const double* __restrict__ pold,
                                                          const double* __restrict__ rho, double* __restrict__ u,
                                                          double* __restrict__ v, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long c = uidx(g, i, j, k), q = pidx(g, i, j, k);
  const long qs[3] = {q + 1, q + g.sp1, q + (long)g.sp1 * g.sp2};
  const double dl[3] = {dxi, dyi, dzci[k]};
  double* vel[3] = {u, v, w};
  const double pc = p[q], rc = rho[q];
  const double e = __dsub_rn(__dmul_rn(f1, pc), __dmul_rn(f2, pold[q]));
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double rhoi = __ddiv_rn(1.0, __dmul_rn(0.5, __dadd_rn(rho[qs[d]], rc)));
    const double A = __dmul_rn(-__dmul_rn(__dsub_rn(p[qs[d]], pc), dl[d]), rho0i);
    const double en = __dsub_rn(__dmul_rn(f1, p[qs[d]]), __dmul_rn(f2, pold[qs[d]]));
    const double B = __dmul_rn(__dmul_rn(__dsub_rn(rhoi, rho0i), __dsub_rn(en, e)), dl[d]);
    vel[d][c] = __dadd_rn(vel[d][c], __dmul_rn(f_t12, __dsub_rn(A, B)));
  }
}

// main__single_phase.f90:693-699 (mode 0: pold = p) and :734-740 (mode 1: p = pold + p), interior only
__global__ void __launch_bounds__(256) pold_update_kernel(StencilGeom g, int mode, double* __restrict__ p, double* __restrict__ pold) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  if (i > g.nx || j > g.ny) return;
  const long q = pidx(g, i, j, k);
  if (mode == 0) pold[q] = p[q];
  else p[q] = __dadd_rn(pold[q], p[q]);
}

// chkdiv.f90:46-58: per-block partial (sum, max|div|), finished by chkdiv_final_kernel (deterministic)
__global__ void __launch_bounds__(256) chkdiv_kernel(StencilGeom g, double dxi, double dyi,
                                                     const double* __restrict__ dzfi, const double* __restrict__ u,
                                                     const double* __restrict__ v, const double* __restrict__ w,
                                                     double* __restrict__ part_sum, double* __restrict__ part_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  double div = 0.0;
  if (i <= g.nx && j <= g.ny) {
    const long c = uidx(g, i, j, k);
    div = __dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(w[c], w[c - g.su1 * g.su2]), dzfi[k]),
                              __dmul_rn(__dsub_rn(v[c], v[c - g.su1]), dyi)),
                    __dmul_rn(__dsub_rn(u[c], u[c - 1]), dxi));
  }
  double s = div, m = fabs(div);
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  }
  __shared__ double ss[8], sm[8];
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if ((t & 31) == 0) { ss[t >> 5] = s; sm[t >> 5] = m; }
  __syncthreads();
  if (t == 0) {
    const int nw = (blockDim.x * blockDim.y + 31) / 32;
    for (int q = 1; q < nw; ++q) { s += ss[q]; m = fmax(m, sm[q]); }
    const long bid = blockIdx.x + (long)gridDim.x * (blockIdx.y + (long)gridDim.y * blockIdx.z);
    part_sum[bid] = s; part_max[bid] = m;
  }
}

__global__ void chkdiv_final_kernel(long nparts, const double* __restrict__ part_sum,
                                    const double* __restrict__ part_max, double* __restrict__ out) {
  __shared__ double ss[256], sm[256];
  double s = 0.0, m = 0.0;
  for (long q = threadIdx.x; q < nparts; q += blockDim.x) { s += part_sum[q]; m = fmax(m, part_max[q]); }
  ss[threadIdx.x] = s; sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = ss[0]; out[1] = sm[0]; }
}

}  // namespace fb

// ------------------------------------------------------------------------------------------------
// boundp (src/bound.f90:146-225): ghost cells of p, halo width 1.  One launch per step of the reference's
// sequence (y halo, z halo, x faces, y faces, z faces) so that edges and corners come out bit-identical.
// A "face pair" in direction d: ghost planes at index 0 and n_d+1, over the full (halo-inclusive) extent
// of the other two directions.  sd = element stride of direction d; (na, sa), (nb, sb) span the face.
struct FaceGeom {
  long sd, sa, sb;
  int nd, na, nb;      // nd = interior size along d; na, nb = full extents (with halos) of the face
};

// periodic wrap (set_bc case 'P', bound.f90:268-318; = updthalo with the rank as its own neighbour)
__global__ void boundp_wrap_kernel(FaceGeom g, double* __restrict__ p) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)g.na * g.nb) return;
  const long b = idx / g.na, a = idx - b * g.na;
  double* q = p + a * g.sa + b * g.sb;
  q[0] = q[(long)g.nd * g.sd];
  q[(long)(g.nd + 1) * g.sd] = q[g.sd];
}

// Dirichlet / Neumann ghost: p_ghost = factor + sgn * p_inner (bound.f90:320-420), side 0 or 1
__global__ void boundp_face_kernel(FaceGeom g, int side, double factor, double sgn, double* __restrict__ p) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)g.na * g.nb) return;
  const long b = idx / g.na, a = idx - b * g.na;
  double* q = p + a * g.sa + b * g.sb;
  if (side == 0) q[0] = __dadd_rn(factor, __dmul_rn(sgn, q[g.sd]));
  else q[(long)(g.nd + 1) * g.sd] = __dadd_rn(factor, __dmul_rn(sgn, q[(long)g.nd * g.sd]));
}

// z halo from neighbouring slabs: contiguous planes (rank-local copies; the inter-GPU transfer itself is the
// caller's exchange callback or a peer copy)
__global__ void copy_plane_kernel(long n, const double* __restrict__ src, double* __restrict__ dst) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) dst[idx] = src[idx];
}
