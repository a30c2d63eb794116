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
#include <cstdint>

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
#include <cuda.h>
#include <cudaTypedefs.h>

namespace fb {

// ---- TMA (cp.async.bulk.tensor) plumbing ---------------------------------------------------------------------
// The 16-column tile is a 2-D box {16 columns, up to 256 levels} of the (ncol x nz) work array: nz/256 bulk tensor copies
// issued by ONE thread bring a whole tile into shared memory and signal an mbarrier, instead of L cp.async per thread
// (ncu, v9 kernel at 1024^3: 17 % issue utilisation, top stalls lg_throttle 5.1 and mio_throttle 4.1 per issue -- the
// LSU queue, not HBM; one LDGSTS costs ~8 issue cycles, 32 per thread x 16 warps ~ 4000 cycles of a 15000-cycle tile).
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// 1-D bulk copy global -> shared, completion counted on the same mbarrier (bytes: multiple of 16, 16-byte aligned both sides)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int L, int TI, int MINB, bool CORR = false>
__global__ void __launch_bounds__(512 / MINB, MINB)
thomas_uni_tma_kernel(long ncol, long ntiles, ThomasArgs T, const double* __restrict__ lam, const __grid_constant__ CUtensorMap tmap,
                      int box_rows, ColGeom og, ThomasCorr corr) {
  using TU = ThomasUni<L, TI>;
  using TR = ThomasReg<L, TI>;
  extern __shared__ __align__(128) double smem[];
  const int S = T.S, nz = T.nz;
  const int tid = threadIdx.x;
  const int st = S * TI;
  double* tile_s = smem;                                  // [nz][TI] dense, filled by the TMA unit
  double* tab = tile_s + (size_t)nz * TI;
  double* ex = tab + TU::tab_doubles();                   // 6 arrays
  double* pcrA = ex + 6 * (size_t)st;
  double* pcrB = pcrA + 3 * (size_t)st;
  double* X = pcrB + 3 * (size_t)st;
  double* lam_s = X + st;                                 // [TI] eigenvalues of the tile in flight, 16-byte aligned
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(lam_s + TI);
  const int lane = tid % TI, s = tid / TI;
  const int k0 = s * L;
  const bool one_chunk = (og.n3l % L) == 0;
  double* obase = nullptr;
  if (one_chunk) { const int q = k0 / og.n3l; obase = og.ptr[q] + og.koff + ncol * (long)(k0 - q * og.n3l); }
  const unsigned tile_bytes = (unsigned)nz * TI * sizeof(double);

  // The tile's eigenvalues ride on the same mbarrier as the tile (one 1-D bulk copy of <= 128 bytes).  r02 source-level
  // profile of the previous version, which prefetched lambda of the NEXT tile into a register with __ldg: the register was
  // spilled, so every thread stalled a full DRAM latency on `LDG -> STL` at the top of each iteration (13.6 % of all
  // stall samples, the largest single site; the mbarrier wait itself: 0 %).
  auto fetch = [&](long tile) {                            // one thread: whole tile, completion counted in bytes on `bar`
    if (tid == 0 && tile < ntiles) {
      const long c0 = tile * TI;
      const unsigned lam_bytes = (unsigned)((ncol - c0 < TI ? ncol - c0 : TI) * sizeof(double));   // ncol is even: multiple of 16
      mbar_expect_tx(bar, tile_bytes + lam_bytes);         // (out-of-range columns of a ragged last tile are zero-filled and counted)
      for (int r = 0; r < nz; r += box_rows) tma_load_2d(tile_s + (size_t)r * TI, &tmap, (int)c0, r, bar);
      bulk_load_1d(lam_s, lam + c0, lam_bytes, bar);
    }
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long tile = blockIdx.x;
  if (tile < ntiles) fetch(tile);

  for (unsigned it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const long col = tile * TI + lane;
    const bool live = col < ncol;
    double v[L];
    mbar_wait(bar, it & 1u);
    const double lm = live ? lam_s[lane] : -1.0;           // dead lanes of a ragged last tile: any regular column
    const bool pin = T.singular && live && (lm == 0.0);
    if (tid < 2 * TI) TU::build(tab, T, lm, lane, tid / TI);
    {
      const double* ts = tile_s + (size_t)k0 * TI + lane;
#pragma unroll
      for (int l = 0; l < L; ++l) v[l] = ts[l * TI];
    }
    double yv[CORR ? L : 1];
    if (CORR) {
#pragma unroll
      for (int l = 0; l < L; ++l) { yv[l] = v[l]; v[l] = 0.0; }
      if (live) {
        if (s == 0 && corr.ca != 0.0) v[0] = -corr.ca * __ldg(corr.xprev + col);
        if (s == S - 1 && corr.cc != 0.0) v[L - 1] = v[L - 1] - corr.cc * __ldg(corr.xnext + col);
      }
    }
    __syncthreads();                                        // tables ready; every thread has its levels: the tile buffer is free
    fetch(tile + gridDim.x);                                // next tile streams in during the rest of this iteration
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
    if (CORR) {
#pragma unroll
      for (int l = 0; l < L; ++l) v[l] += yv[l];
    }
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

// General (coefficient-table or scalar-coefficient) register kernel of thomas_reg.cuh with the tile brought in by TMA:
// same phases as thomas_reg_kernel<.., CL = 1>, the per-thread cp.async slots replaced by one dense [nz][TI] tile.
template <int L, int TI, bool UNI>
__global__ void __launch_bounds__(512, 1)
thomas_reg_tma_kernel(long ncol, long ntiles, ThomasArgs T, const double* __restrict__ lam, const __grid_constant__ CUtensorMap tmap,
                      int box_rows, ColGeom og) {
  using TR = ThomasReg<L, TI>;
  extern __shared__ __align__(128) double smem[];
  const int nz = T.nz, S = T.S;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int st = S * TI;
  double* tile_s = smem;                                  // [nz][TI] dense, filled by the TMA unit
  double* ex = tile_s + (size_t)nz * TI;                  // 6 arrays
  double* pcrA = ex + 6 * (size_t)st;
  double* pcrB = pcrA + 3 * (size_t)st;
  double* X = pcrB + 3 * (size_t)st;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(X + st);
  double* coef = X + st + 2;                              // az | bz | cz at padded rows
  const int lane = tid % TI, s = tid / TI;
  if (!UNI) {
    const int tr = TR::tile_rows(nz);
    for (int k = tid; k < nz; k += nthr) {
      const int r = TR::prow(k);
      coef[r] = __ldg(T.az + k); coef[tr + r] = __ldg(T.bz + k); coef[2 * tr + r] = __ldg(T.cz + k);
    }
    T.az = coef; T.bz = coef + tr; T.cz = coef + 2 * tr; T.padded = 1;
  }
  using CF = typename std::conditional<UNI, CoefUniform<L>, CoefTable<L>>::type;
  const int k0 = s * L;
  const bool one_chunk = (og.n3l % L) == 0;
  double* obase = nullptr;
  if (one_chunk) { const int q = k0 / og.n3l; obase = og.ptr[q] + og.koff + ncol * (long)(k0 - q * og.n3l); }
  const unsigned tile_bytes = (unsigned)nz * TI * sizeof(double);
  auto fetch = [&](long tile) {
    if (tid == 0 && tile < ntiles) {
      mbar_expect_tx(bar, tile_bytes);
      for (int r = 0; r < nz; r += box_rows) tma_load_2d(tile_s + (size_t)r * TI, &tmap, (int)(tile * TI), r, bar);
    }
  };
  auto lam_of = [&](long tile) {
    const long col = tile * TI + lane;
    return (col < ncol) ? __ldg(lam + col) : -1.0;
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                        // barrier initialised, coefficients staged
  long tile = blockIdx.x;
  double lm_next = 0.0;
  if (tile < ntiles) { fetch(tile); lm_next = lam_of(tile); }

  for (unsigned it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const long col = tile * TI + lane;
    const bool live = col < ncol;
    const double lm = lm_next;
    if (tile + gridDim.x < ntiles) lm_next = lam_of(tile + gridDim.x);
    const bool pin = T.singular && live && (lm == 0.0);
    double v[L];
    mbar_wait(bar, it & 1u);
    {
      const double* ts = tile_s + (size_t)k0 * TI + lane;
#pragma unroll
      for (int l = 0; l < L; ++l) v[l] = ts[l * TI];
    }
    const CF cf(T, s);
    SegRegs<L> g;
    TR::phase1(v, T, cf, lm, lane, s, g, ex);
    __syncthreads();                                        // every thread holds its levels: the tile buffer is free
    fetch(tile + gridDim.x);
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
    TR::phase3(v, X, T, lane, s, g);
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
    // X / ex / pcr buffers are rewritten only after the next iteration's barriers
  }
}

// tensor map of the (ncol x nz) FP64 work array with a {TI, box_rows} box; cached for the last (pointer, shape)
inline cudaError_t thomas_uni_tensor_map(const double* W, long ncol, int nz, int ti, int box_rows, CUtensorMap* out) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  static const double* cW = nullptr; static long cncol = 0; static int cnz = 0, cti = 0, cbr = 0; static CUtensorMap cmap;
  if (cW != W || cncol != ncol || cnz != nz || cti != ti || cbr != box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)ncol, (cuuint64_t)nz};
    const cuuint64_t gstr[1] = {(cuuint64_t)ncol * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)ti, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    static const int promo = [] { const char* e = getenv("FLUTAS_B200_TMA_L2"); return e ? atoi(e) : 128; }();
    const CUtensorMapL2promotion l2 = promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                    : promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    const CUresult r = encode(&cmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(W), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cW = nullptr; return cudaErrorInvalidValue; }
    cW = W; cncol = ncol; cnz = nz; cti = ti; cbr = box_rows;
  }
  *out = cmap;
  return cudaSuccess;
}

inline int threads_of(const ThomasArgs& T, int ti) { return ti * T.S; }

template <int L, int TI, int MINB, bool CORR = false>
inline cudaError_t thomas_uni_tma_launch(long ncol, const ThomasArgs& T, const double* lam, const double* W, const ColGeom& og,
                                         int nsm, cudaStream_t st, const ThomasCorr* corr = nullptr) {
  using TU = ThomasUni<L, TI>;
  auto kern = thomas_uni_tma_kernel<L, TI, MINB, CORR>;
  if (threads_of(T, TI) > 512 / MINB) return cudaErrorInvalidValue;
  const ThomasCorr cr = corr ? *corr : ThomasCorr{nullptr, nullptr, 0.0, 0.0};
  const int threads = TI * T.S;
  const int box_rows = T.nz < 256 ? T.nz : 256;
  if (T.nz % box_rows) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)T.nz * TI + TU::tab_doubles() + 13 * (size_t)T.S * TI + TI + 2) * sizeof(double);
  const long ntiles = (ncol + TI - 1) / TI;
  CUtensorMap map;
  cudaError_t e = thomas_uni_tensor_map(W, ncol, T.nz, TI, box_rows, &map);
  if (e != cudaSuccess) return e;
  static int per_sm = 0, cfg_nz = 0;
  if (per_sm == 0 || cfg_nz != T.nz) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (q < 1) return cudaErrorLaunchOutOfResources;
    per_sm = q; cfg_nz = T.nz;
  }
  const long grid = ntiles < (long)nsm * per_sm ? ntiles : (long)nsm * per_sm;
  kern<<<(unsigned)grid, threads, smem, st>>>(ncol, ntiles, T, lam, map, box_rows, og, cr);
  return cudaGetLastError();
}

template <int L, int TI, bool UNI>
inline cudaError_t thomas_reg_tma_launch(long ncol, const ThomasArgs& T, const double* lam, const double* W, const ColGeom& og,
                                         int nsm, cudaStream_t st) {
  using TR = ThomasReg<L, TI>;
  auto kern = thomas_reg_tma_kernel<L, TI, UNI>;
  const int threads = TI * T.S;
  const int box_rows = T.nz < 256 ? T.nz : 256;
  if (T.nz % box_rows || threads > 512) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)T.nz * TI + 13 * (size_t)T.S * TI + 2 + (UNI ? 0 : 3 * (size_t)TR::tile_rows(T.nz))) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const long ntiles = (ncol + TI - 1) / TI;
  CUtensorMap map;
  cudaError_t e = thomas_uni_tensor_map(W, ncol, T.nz, TI, box_rows, &map);
  if (e != cudaSuccess) return e;
  static int per_sm = 0, cfg_nz = 0;
  if (per_sm == 0 || cfg_nz != T.nz) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (q < 1) return cudaErrorLaunchOutOfResources;
    per_sm = q; cfg_nz = T.nz;
  }
  const long grid = ntiles < (long)nsm * per_sm ? ntiles : (long)nsm * per_sm;
  kern<<<(unsigned)grid, threads, smem, st>>>(ncol, ntiles, T, lam, map, box_rows, og);
  return cudaGetLastError();
}

// General kernel with TMA tile loads: 16-column tiles when a column has <= 32 segments, else 8-column tiles (<= 64).
// *done = false when the shape / alignment is not served (caller: thomas_reg_run with cp.async slots).
inline int thomas_reg_tma_run(long ncol, int nz, const double* az, const double* bz, const double* cz, const double* lam,
                              const double* W, double* Wout, const ColGeom* out, bool periodic, int singular, int nsm,
                              const ThomasArgs* uni, cudaStream_t st, bool* done) {
  *done = false;
  // Measured on B200: no gain over the cp.async slots for this kernel (512^3 0.492 vs 0.493 ms; 8-column tiles at 1024^3
  // 7.2 vs 6.4 ms) -- its 16 copies per thread do not saturate the LSU queue the way the 32 of the shared-LU kernel do.
  // Kept as an option: FLUTAS_B200_THOMAS_TMA_GEN=1.
  static const bool tma_env = [] { const char* e = getenv("FLUTAS_B200_THOMAS_TMA_GEN"); return e && e[0] == '1'; }();
  int L = 0;
  if (!tma_env || (ncol % 2) || (reinterpret_cast<uintptr_t>(W) % 16) || !thomas_reg_pick(nz, periodic, &L) || L != 16) return 0;
  ThomasArgs T;
  T.nz = nz; T.S = nz / L; T.periodic = periodic ? 1 : 0; T.singular = singular; T.az = az; T.bz = bz; T.cz = cz;
  T.padded = 0; T.uniform = 0;
  if (uni && uni->uniform) {
    T.uniform = 1; T.a0 = uni->a0; T.b0 = uni->b0; T.a_first = uni->a_first; T.b_first = uni->b_first;
    T.b_last = uni->b_last; T.c_last = uni->c_last;
  }
  ColGeom og;
  if (out) og = *out;
  else { for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = Wout; og.n3l = nz; og.koff = 0; }
  cudaError_t e;
  if (T.S <= 32) e = T.uniform ? thomas_reg_tma_launch<16, 16, true>(ncol, T, lam, W, og, nsm, st) : thomas_reg_tma_launch<16, 16, false>(ncol, T, lam, W, og, nsm, st);
  else if (T.S <= 64) e = T.uniform ? thomas_reg_tma_launch<16, 8, true>(ncol, T, lam, W, og, nsm, st) : thomas_reg_tma_launch<16, 8, false>(ncol, T, lam, W, og, nsm, st);
  else return 0;
  if (e == cudaErrorInvalidValue) { (void)cudaGetLastError(); return 0; }     // shape not served (shared memory): fall back
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
}

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

// Rank-local block of the distributed z solve: nz = n3l levels, never periodic (the wrap-around coupling lives in the
// interface system), L = 16 levels per thread, S = nz/16 in 2..32 segments -> 32 S threads per tile, 4 / 2 / 1 blocks per SM.
// corr = nullptr: pass 1 (y = T_g^{-1} b); else pass 2 (x = y + correction).  *done = false: shape not served.
inline bool thomas_uni_local_ok(int nz, long ncol, const void* W, const void* lam) {
  return nz % 16 == 0 && nz / 16 >= 2 && nz / 16 <= 32 && (ncol % 2) == 0 && (reinterpret_cast<uintptr_t>(W) % 16) == 0 &&
         (reinterpret_cast<uintptr_t>(lam) % 16) == 0;
}
inline int thomas_uni_local_run(long ncol, int nz, const double* lam, double* W, int singular, int nsm, const ThomasArgs* uni,
                                const ThomasCorr* corr, cudaStream_t st, bool* done) {
  *done = false;
  if (!uni || !uni->uniform || !thomas_uni_local_ok(nz, ncol, W, lam)) return 0;
  ThomasArgs T = *uni;
  T.nz = nz; T.S = nz / 16; T.periodic = 0; T.singular = singular; T.az = T.bz = T.cz = nullptr; T.padded = 0;
  ColGeom og;
  for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = W;
  og.n3l = nz; og.koff = 0;
  cudaError_t e;
  if (T.S <= 8) e = corr ? thomas_uni_tma_launch<16, 16, 4, true>(ncol, T, lam, W, og, nsm, st, corr) : thomas_uni_tma_launch<16, 16, 4, false>(ncol, T, lam, W, og, nsm, st);
  else if (T.S <= 16) e = corr ? thomas_uni_tma_launch<16, 16, 2, true>(ncol, T, lam, W, og, nsm, st, corr) : thomas_uni_tma_launch<16, 16, 2, false>(ncol, T, lam, W, og, nsm, st);
  else e = corr ? thomas_uni_tma_launch<16, 16, 1, true>(ncol, T, lam, W, og, nsm, st, corr) : thomas_uni_tma_launch<16, 16, 1, false>(ncol, T, lam, W, og, nsm, st);
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
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
  // TMA tile loads need a 16-byte aligned base and row pitch (ncol even); FLUTAS_B200_THOMAS_TMA=0 keeps the cp.async kernel
  static const bool tma_env = [] { const char* e = getenv("FLUTAS_B200_THOMAS_TMA"); return !(e && e[0] == '0'); }();
  if (tma_env && (ncol % 2) == 0 && (reinterpret_cast<uintptr_t>(W) % 16) == 0 && (reinterpret_cast<uintptr_t>(lam) % 16) == 0) {
    switch (L) {
      case 4: e = thomas_uni_tma_launch<4, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
      case 8: e = thomas_uni_tma_launch<8, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
      case 16: e = thomas_uni_tma_launch<16, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
      default: e = thomas_uni_tma_launch<32, 16, 1>(ncol, T, lam, W, og, nsm, st); break;
    }
    if (e != cudaSuccess) return (int)e;
    *done = true;
    return 0;
  }
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
