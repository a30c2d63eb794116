// thomas_uni.cuh -- z-direction tridiagonal solve on an EXACTLY UNIFORM z grid: the LU factors are shared per column.
//
// Same two-level partition method as thomas_reg.cuh (segments of L levels per thread, reduced system in the
// separators solved by PCR; replaces gaussel / gaussel_periodic, src/solver_cpu.f90:117-185).  On a uniform grid
// every interior row of the matrix is (a0, b0 + lambda, a0): all segments of a column share ONE factorisation
// (z_l, d_l = a0 z_l, f_l and the couplings F0, G0 of the first interior row), only the first segment differs (its
// first row carries the boundary condition).  So per 16-column tile 32 threads build two small tables in shared
// memory and the other threads only sweep the right-hand side:
//     forward  r_l = v_l z_l - d_l r_{l-1}          2 FP64 instructions per level
//     R0       R   = r_l - d_l R                    1
//     final    x_l = (r_l - f_l X_{s-1}) - d_l x_{l+1}   2
// against 19 in the general kernel, and the per-thread state shrinks from 3 L + L to L doubles.  That is what lets
// one SM hold a 16-COLUMN tile of nz = 1024 levels (L = 32, 512 threads): 128-byte row pieces instead of 64-byte
// ones.  tools/pattern_bench.cu (same access pattern, no arithmetic): 8-column tiles at 1024^3 top out at 2.6 TB/s
// -- exactly what the general kernel reaches there -- 16-column tiles at 5.1 TB/s.
//
// Which grids: initgrid.f90 with gr = 0 gives bit-identical a, b, c rows whenever lz/nz is exactly representable
// (every BASELINE channel/RB configuration: lz = 1, nz a power of two); otherwise (lz = 2 pi) the rows differ in
// the last bit and the general kernel (coefficient tables) is used -- never an averaged coefficient.
// Host-compilable core (tests/emulate), like thomas_reg.cuh.
#pragma once
#include "thomas_reg.cuh"

namespace fb {

template <int L, int TI>
struct ThomasUni {
  // tables (doubles): [variant 0 = first segment, 1 = the others][array z | d | f][level 0..L-2 (L rows kept)][TI]
  // followed by the per-column couplings [variant][F0 | G0][TI]
  static FB_HD size_t tab_doubles() { return (size_t)2 * 3 * L * TI + (size_t)2 * 2 * TI; }
  static FB_HD double* tz(double* tab, int var) { return tab + (size_t)var * 3 * L * TI; }
  static FB_HD double* td(double* tab, int var) { return tz(tab, var) + (size_t)L * TI; }
  static FB_HD double* tf(double* tab, int var) { return tz(tab, var) + (size_t)2 * L * TI; }
  static FB_HD double* tF0(double* tab, int var) { return tab + (size_t)2 * 3 * L * TI + (size_t)var * 2 * TI; }
  static FB_HD double* tG0(double* tab, int var) { return tF0(tab, var) + TI; }

  // One thread per (variant, column): factorise the L-1 interior rows of a segment.
  // Pivots through the leading principal minors of the matrix scaled by 1/a0 (off-diagonals 1, diagonal
  // beta = (b0 + lambda)/a0): t_l = beta_l t_{l-1} - g_l t_{l-2}, one dependent FMA per level, then independent
  // reciprocals; z_l = t_{l-1} / (a0 t_l).  |t_l| <= (2 + |lambda|/a0)^l stays far inside the double range for L <= 32.
  static FB_HD void build(double* tab, const ThomasArgs& T, double lam, int lane, int var) {
    const double a0 = T.a0, ia0 = fb_rcp(a0);
    const bool first = (var == 0);
    const double afirst = first ? T.a_first : a0;                        // sub-diagonal of row 0 (couples to X_{s-1})
    const double beta0 = ((first ? T.b_first : T.b0) + lam) * ia0, beta = (T.b0 + lam) * ia0;
    double* z = tz(tab, var) + lane; double* d = td(tab, var) + lane; double* f = tf(tab, var) + lane;
    double dr[L], fr[L];                                                  // kept in registers for the back substitution
    double tm = 1.0, t = beta0;
    double zl = fb_rcp(t) * ia0;                                          // z_0 = 1 / (b_0 + lambda)
    dr[0] = a0 * zl; fr[0] = afirst * zl;
    z[0] = zl; d[0] = dr[0]; f[0] = fr[0];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = 1; l < L - 1; ++l) {
      const double tn = beta * t - tm;                                    // a_l c_{l-1} / a0^2 = 1
      zl = t * fb_rcp(tn) * ia0;
      dr[l] = a0 * zl;
      fr[l] = -dr[l] * fr[l - 1];
      z[l * TI] = zl; d[l * TI] = dr[l]; f[l * TI] = fr[l];
      tm = t; t = tn;
    }
    // first interior row in terms of the separators: x_0 = R0 - F0 X_{s-1} - G0 X_s
    double F = fr[L - 2], G = dr[L - 2];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = L - 3; l >= 0; --l) {
      F = fr[l] - dr[l] * F;
      G = -dr[l] * G;
    }
    tF0(tab, var)[lane] = F; tG0(tab, var)[lane] = G;
  }

  // forward sweep of this thread's right-hand side (in place: v_l <- r_l for l <= L-2, v_{L-1} untouched), then the
  // six reduced-system inputs of segment s in the layout ThomasReg::reduced_row reads
  static FB_HD void phase1(double* v, const double* tab, const ThomasArgs& T, int lane, int s, double* ex) {
    const int var = (s == 0) ? 0 : 1;
    const double* z = tz(const_cast<double*>(tab), var) + lane;
    const double* d = td(const_cast<double*>(tab), var) + lane;
    double r = v[0] * z[0];
    v[0] = r;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = 1; l < L - 1; ++l) { r = v[l] * z[l * TI] - d[l * TI] * r; v[l] = r; }
    double R = r;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = L - 3; l >= 0; --l) R = v[l] - d[l * TI] * R;
    const int o = s * TI + lane, st = T.S * TI;
    ex[o] = R;
    ex[st + o] = tF0(const_cast<double*>(tab), var)[lane];
    ex[2 * st + o] = tG0(const_cast<double*>(tab), var)[lane];
    ex[3 * st + o] = r;
    ex[4 * st + o] = tf(const_cast<double*>(tab), var)[(L - 2) * TI + lane];
    ex[5 * st + o] = d[(L - 2) * TI];
  }

  // substitution with the known separators
  static FB_HD void phase3(double* v, const double* X, const double* tab, const ThomasArgs& T, int lane, int s) {
    const int S = T.S, var = (s == 0) ? 0 : 1;
    const double* d = td(const_cast<double*>(tab), var) + lane;
    const double* f = tf(const_cast<double*>(tab), var) + lane;
    const double xs = X[s * TI + lane];
    const double xp = X[((s == 0) ? S - 1 : s - 1) * TI + lane];      // multiplied by f = 0 in segment 0 unless periodic
    double x = xs;
    v[L - 1] = xs;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = L - 2; l >= 0; --l) { x = (v[l] - f[l * TI] * xp) - d[l * TI] * x; v[l] = x; }
  }
};

// segment length of the uniform kernel: S = nz / L <= 32 segments (512 threads at 16 columns), L in {4, 8, 16, 32}
inline bool thomas_uni_pick(int nz, bool periodic, int* Lout) {
  const int cand[4] = {4, 8, 16, 32};
  for (int q = 0; q < 4; ++q) {
    const int L = cand[q];
    if (nz % L) continue;
    const int S = nz / L;
    if (S < 2 || S > 32) continue;
    if (periodic && (S & (S - 1))) continue;
    *Lout = L;
    return true;
  }
  return false;
}

}  // namespace fb

#if defined(__CUDACC__)
namespace fb {

template <int L, int TI, int MINB>
__global__ void __launch_bounds__(512 / MINB, MINB)
thomas_uni_kernel(long ncol, long ntiles, ThomasArgs T, const double* __restrict__ lam, const double* W, ColGeom og) {
  using TU = ThomasUni<L, TI>;
  using TR = ThomasReg<L, TI>;
  constexpr int MAXT = 512 / MINB;
  extern __shared__ double smem[];
  const int S = T.S;
  const int tid = threadIdx.x;
  const int st = S * TI;
  double* slots = smem;                                   // [L][MAXT], private per thread
  double* tab = slots + (size_t)L * MAXT;
  double* ex = tab + TU::tab_doubles();                   // 6 arrays
  double* pcrA = ex + 6 * (size_t)st;
  double* pcrB = pcrA + 3 * (size_t)st;
  double* X = pcrB + 3 * (size_t)st;
  const int lane = tid % TI, s = tid / TI;
  const int k0 = s * L;
  const bool one_chunk = (og.n3l % L) == 0;
  double* obase = nullptr;
  if (one_chunk) { const int q = k0 / og.n3l; obase = og.ptr[q] + og.koff + ncol * (long)(k0 - q * og.n3l); }

  auto fetch = [&](long tile) {
    if (tile < ntiles) {
      const long col = min(tile * TI + lane, ncol - 1);
      const double* src = W + col + (long)k0 * ncol;
      double* sl = slots + tid;
#pragma unroll
      for (int l = 0; l < L; ++l) cp_async8(sl + l * MAXT, src + (long)l * ncol);
    }
    cp_async_commit();
  };
  auto lam_of = [&](long tile) {                           // dead lanes of a ragged last tile: any regular column
    const long col = tile * TI + lane;
    return (col < ncol) ? __ldg(lam + col) : -1.0;
  };
  long tile = blockIdx.x;
  double lm_next = 0.0;
  if (tile < ntiles) { fetch(tile); lm_next = lam_of(tile); }

  for (; tile < ntiles; tile += gridDim.x) {
    const long col = tile * TI + lane;
    const bool live = col < ncol;
    const double lm = lm_next;
    if (tile + gridDim.x < ntiles) lm_next = lam_of(tile + gridDim.x);
    const bool pin = T.singular && live && (lm == 0.0);
    if (tid < 2 * TI) TU::build(tab, T, lm, lane, tid / TI);           // threads (s = 0, 1) hold every column's lambda
    double v[L];
    cp_async_wait_all();
#pragma unroll
    for (int l = 0; l < L; ++l) v[l] = slots[l * MAXT + tid];
    fetch(tile + gridDim.x);                                // the slots are private per thread and empty again: the next
                                                            // tile streams in during the whole of this iteration
    __syncthreads();                                        // tables ready (the previous tile's readers left at its last barrier)
    TU::phase1(v, tab, T, lane, s, ex);
    __syncthreads();
    const CoefUniform<L> cf(T, s);
    TR::reduced_row(v[L - 1], ex, pcrA, T, cf, lm, lane, s, pin);
    __syncthreads();
    double* src = pcrA;
    double* dst = pcrB;
    const int hmax = T.periodic ? S / 2 : S;
    for (int h = 1; h < hmax; h *= 2) {
      const bool coupled = TR::pcr_step(src, dst, T, lane, s, h);
      const int any = __syncthreads_or(coupled ? 1 : 0);
      double* t = src; src = dst; dst = t;
      if (!any) break;
    }
    TR::pcr_finish(src, X, T, lane, s);
    __syncthreads();
    TU::phase3(v, X, tab, T, lane, s);
    if (live) {
      if (one_chunk) {
        double* dstp = obase + col;
#pragma unroll
        for (int l = 0; l < L; ++l) st_z(dstp + (long)l * ncol, v[l]);
      } else {
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const int k = k0 + l, q = k / og.n3l;
          st_z(og.ptr[q] + og.koff + col + ncol * (long)(k - q * og.n3l), v[l]);
        }
      }
    }
    __syncthreads();                                        // the tables are rebuilt by the next iteration
  }
}

template <int L, int TI, int MINB>
inline cudaError_t thomas_uni_launch(long ncol, const ThomasArgs& T, const double* lam, const double* W, const ColGeom& og,
                                     int nsm, cudaStream_t st) {
  using TU = ThomasUni<L, TI>;
  auto kern = thomas_uni_kernel<L, TI, MINB>;
  constexpr int MAXT = 512 / MINB;
  const int threads = TI * T.S;
  const size_t smem = ((size_t)L * MAXT + TU::tab_doubles() + 13 * (size_t)T.S * TI) * sizeof(double);
  const long ntiles = (ncol + TI - 1) / TI;
  static int per_sm = 0, cfg_nz = 0;
  if (per_sm == 0 || cfg_nz != T.nz) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (q < 1) return cudaErrorLaunchOutOfResources;
    per_sm = q; cfg_nz = T.nz;
  }
  const long grid = ntiles < (long)nsm * per_sm ? ntiles : (long)nsm * per_sm;
  kern<<<(unsigned)grid, threads, smem, st>>>(ncol, ntiles, T, lam, W, og);
  return cudaGetLastError();
}

// *done = false if the grid is not exactly uniform or nz is not served (caller: thomas_reg_run).
inline int thomas_uni_run(long ncol, int nz, const double* lam, const double* W, double* Wout, const ColGeom* out, bool periodic,
                          int singular, int nsm, const ThomasArgs* uni, cudaStream_t st, bool* done) {
  *done = false;
  int L = 0;
  if (!uni || !uni->uniform || !thomas_uni_pick(nz, periodic, &L)) return 0;
  ThomasArgs T = *uni;
  T.nz = nz; T.S = nz / L; T.periodic = periodic ? 1 : 0; T.singular = singular; T.az = T.bz = T.cz = nullptr; T.padded = 0;
  ColGeom og;
  if (out) og = *out;
  else { for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = Wout; og.n3l = nz; og.koff = 0; }
  cudaError_t e = cudaSuccess;
  switch (L) {
    case 4: e = thomas_uni_launch<4, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
    case 8: e = thomas_uni_launch<8, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
    case 16: e = thomas_uni_launch<16, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
    default: e = thomas_uni_launch<32, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
  }
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
}

}  // namespace fb
#endif
