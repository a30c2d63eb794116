// thomas_reg.cuh -- z-direction tridiagonal solve with every column segment held in REGISTERS.
//
// Replaces gaussel / gaussel_periodic (src/solver_cpu.f90:117-185) and the reference GPU versions that
// spill 2-4 scratch fields to memory (src/solver_gpu.f90:475-638).  Same two-level partition method as
// thomas_tile.cuh (that kernel stays as the fall-back for shapes this one does not serve), re-organised
// after the round-1 profile (3.9 warp-instructions per point, 31 % of the HBM peak):
//
//   * thread (lane, s) owns levels [sL, (s+1)L) of column `lane` of a TI-column tile and keeps them in
//     registers from the global load to the global store -- the field never sits in a shared-memory tile;
//   * ONE downward LU sweep per interior row (one reciprocal) that also carries the fill-in f_l towards the
//     previous separator, then a division-free 3-term back substitution that yields the first interior
//     row in terms of the two separators (the UL sweep of thomas_tile.cuh and its second reciprocal go);
//   * (r'_l, f_l, d_l) stay in registers, so the final substitution is two FMAs per level;
//   * the reduced (cyclic) tridiagonal system in the S separators is normalised to a unit diagonal and
//     solved by parallel cyclic reduction in shared memory: one reciprocal per step instead of two;
//   * persistent blocks: the next tile is fetched with cp.async into per-thread private shared-memory
//     slots while the current one is being solved, so loads, arithmetic and stores of one block overlap.
//
// HBM traffic: 16 B/pt (read once, write once).  Host-compilable core (tests/emulate) like tile_fft.cuh.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "thomas_tile.cuh"

namespace fb {

template <int L>
struct SegRegs {                 // per-thread state of the interior rows 0..L-2
  double rp[L], f[L], d[L];
};

// coefficients of this thread's rows l = 0..L-1: from the (padded, shared-memory) tables, or from a handful of
// scalars when the z grid is uniform (no loads, and c_l z_l == a_l z_l is computed once)
template <int L>
struct CoefTable {
  const double *az, *bz, *cz;
  FB_HD CoefTable(const ThomasArgs& T, int s) : az(T.az + s * (L + 1)), bz(T.bz + s * (L + 1)), cz(T.cz + s * (L + 1)) {}
  FB_HD double a(int l) const { return az[l]; }
  FB_HD double b(int l) const { return bz[l]; }
  FB_HD double c(int l) const { return cz[l]; }
};
template <int L>
struct CoefUniform {
  double a0, b0, af, bf, bl, cl;
  bool first, last;
  FB_HD CoefUniform(const ThomasArgs& T, int s)
      : a0(T.a0), b0(T.b0), af(T.a_first), bf(T.b_first), bl(T.b_last), cl(T.c_last), first(s == 0), last(s == T.S - 1) {}
  FB_HD double a(int l) const { return (l == 0 && first) ? af : a0; }
  FB_HD double b(int l) const { return (l == 0 && first) ? bf : (l == L - 1 && last) ? bl : b0; }
  FB_HD double c(int l) const { return (l == L - 1 && last) ? cl : a0; }
};

// true if (a,b,c) describe a uniform grid; fills the scalar fields of T.  az/cz conventions as ThomasArgs.
inline bool thomas_detect_uniform(int nz, const double* a, const double* b, const double* c, bool periodic, ThomasArgs& T) {
  T.uniform = 0; T.a0 = T.b0 = T.a_first = T.b_first = T.b_last = T.c_last = 0.0;
  if (nz < 4) return false;
  const double a0 = a[1], b0 = b[1];
  if (c[0] != a0) return false;
  for (int k = 1; k < nz; ++k) if (a[k] != a0) return false;
  for (int k = 0; k < nz - 1; ++k) if (c[k] != a0) return false;
  for (int k = 1; k < nz - 1; ++k) if (b[k] != b0) return false;
  T.uniform = 1; T.a0 = a0; T.b0 = b0;
  T.a_first = periodic ? a[0] : 0.0; T.b_first = b[0];
  T.b_last = b[nz - 1]; T.c_last = periodic ? c[nz - 1] : 0.0;
  return true;
}

template <int L, int TI>
struct ThomasReg {
  static FB_HD int prow(int k) { return k + k / L; }
  static FB_HD int tile_rows(int nz) { return nz + nz / L; }
  // shared memory (doubles): nbuf fetch slots of L*maxt | ex 6*S*TI | pcr 2*3*S*TI | X S*TI | coefficients 3*tile_rows
  static FB_HD size_t smem_doubles(int nz, int maxt, int nbuf = 1) {
    const size_t S = (size_t)(nz / L), st = S * TI;
    return (size_t)nbuf * L * maxt + 13 * st + 3 * (size_t)tile_rows(nz);
  }

  // ---- phase 1: LU sweep down the interior rows + 3-term back substitution.
  //   x_l = rp_l - f_l X_{s-1} - d_l x_{l+1}   (x_{L-1} = X_s)
  //   ex[0..2] <- (R0, F0, G0): x_0     = R0 - F0 X_{s-1} - G0 X_s
  //   ex[3..5] <- (RD, FD, DD): x_{L-2} = RD - FD X_{s-1} - DD X_s
  template <class CF>
  static FB_HD void phase1(const double* v, const ThomasArgs& T, const CF& cf, double lam, int lane, int s, SegRegs<L>& g,
                           double* ex) {
    // Pivots without a serial division chain: the leading principal minors th_l = bb_l th_{l-1} - a_l c_{l-1} th_{l-2}
    // cost one dependent FMA per level; z_l = 1/(bb_l - a_l c_{l-1} z_{l-1}) = th_{l-1} / th_l are then L-1
    // independent reciprocals (same LU, the quotients are just formed at the end).
    double z[L];
    {
      double thm = 1.0, th = cf.b(0) + lam;
      z[0] = fb_rcp(th);
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int l = 1; l < L - 1; ++l) {
        const double gk = cf.a(l) * cf.c(l - 1);
        const double tn = (cf.b(l) + lam) * th - gk * thm;
        z[l] = th * fb_rcp(tn);
        thm = th; th = tn;
      }
    }
    double rprev = 0.0, fprev = 0.0, dprev = 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = 0; l < L - 1; ++l) {
      const double zz = z[l];
      const double azz = cf.a(l) * zz;
      rprev = v[l] * zz - azz * rprev;
      fprev = (l == 0) ? azz : -azz * fprev;
      dprev = cf.c(l) * zz;
      g.rp[l] = rprev; g.f[l] = fprev; g.d[l] = dprev;
    }
    double R = rprev, F = fprev, G = dprev;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = L - 3; l >= 0; --l) {
      R = g.rp[l] - g.d[l] * R;
      F = g.f[l] - g.d[l] * F;
      G = -g.d[l] * G;
    }
    const int o = s * TI + lane, st = T.S * TI;
    ex[o] = R; ex[st + o] = F; ex[2 * st + o] = G;
    ex[3 * st + o] = rprev; ex[4 * st + o] = fprev; ex[5 * st + o] = dprev;
  }

  // ---- phase 2a: row of separator s of the reduced system, normalised to a unit diagonal: (A, C, R)
  template <class CF>
  static FB_HD void reduced_row(double vsep, const double* ex, double* pcr, const ThomasArgs& T, const CF& cf, double lam,
                                int lane, int s, bool pin) {
    const int S = T.S, st = S * TI, o = s * TI + lane;
    const int sn = (s + 1 == S) ? 0 : s + 1, on = sn * TI + lane;
    const double ak = cf.a(L - 1), ck = cf.c(L - 1), bk = cf.b(L - 1) + lam;   // ck = 0 on the last row unless periodic
    const double r0 = ex[on], f0 = ex[st + on], g0 = ex[2 * st + on];
    const double rd = ex[3 * st + o], fd = ex[4 * st + o], dd = ex[5 * st + o];
    double A = -ak * fd;
    double B = bk - ak * dd - ck * f0;
    double C = -ck * g0;
    double R = vsep - ak * rd - ck * r0;
    if (pin && s == S - 1) { A = 0.0; B = 1.0; C = 0.0; R = 0.0; }        // gauge: x(nz) = 0
    const double inv = fb_rcp(B);
    pcr[o] = A * inv; pcr[st + o] = C * inv; pcr[2 * st + o] = R * inv;
  }

  // ---- phase 2b: one PCR step with stride h on unit-diagonal rows.  Out-of-range neighbours are clamped:
  // their coupling coefficient is already zero in a non-periodic system.
  // Returns true while this row is still coupled to its neighbours.  For a diagonally dominant system the couplings
  // decay doubly exponentially with the step (most (kx,ky) columns are decoupled after 2-3 steps), so the caller
  // stops as soon as no row of the tile reports a coupling above 2^-60 (unit diagonal): dropping such a term changes
  // the solution far below rounding.
  static FB_HD bool pcr_step(const double* src, double* dst, const ThomasArgs& T, int lane, int s, int h) {
    const int S = T.S, st = S * TI, o = s * TI + lane;
    int sm = s - h, sp = s + h;
    if (T.periodic) { sm &= (S - 1); sp &= (S - 1); }
    else { sm = sm < 0 ? 0 : sm; sp = sp >= S ? S - 1 : sp; }
    const int qm = sm * TI + lane, qp = sp * TI + lane;
    const double A = src[o], C = src[st + o], R = src[2 * st + o];
    const double Am = src[qm], Cm = src[st + qm], Rm = src[2 * st + qm];
    const double Ap = src[qp], Cp = src[st + qp], Rp = src[2 * st + qp];
    const double inv = fb_rcp(1.0 - A * Cm - C * Ap);
    const double An = -A * Am * inv, Cn = -C * Cp * inv;
    dst[o] = An;
    dst[st + o] = Cn;
    dst[2 * st + o] = (R - A * Rm - C * Rp) * inv;
    const double tiny = 8.6736173798840355e-19;                            // 2^-60
    return fabs(An) > tiny || fabs(Cn) > tiny;
  }

  // ---- phase 2c: rows are decoupled (non-periodic) or coupled only to row s + S/2 (periodic)
  static FB_HD void pcr_finish(const double* src, double* X, const ThomasArgs& T, int lane, int s) {
    const int S = T.S, st = S * TI, o = s * TI + lane;
    const double R = src[2 * st + o];
    if (T.periodic) {
      const int t = (s + S / 2) & (S - 1), q = t * TI + lane;
      const double K = src[o] + src[st + o], Kt = src[q] + src[st + q], Rt = src[2 * st + q];
      X[o] = (R - K * Rt) * fb_rcp(1.0 - K * Kt);
    } else {
      X[o] = R;
    }
  }

  // ---- phase 3: substitution with the known separators; v <- solution of this thread's L levels
  static FB_HD void phase3(double* v, const double* X, const ThomasArgs& T, int lane, int s, const SegRegs<L>& g) {
    const int S = T.S;
    const int spv = (s == 0) ? S - 1 : s - 1;
    const double xs = X[s * TI + lane];
    const double xp = X[spv * TI + lane];                                  // multiplied by f = 0 in segment 0 unless periodic
    double x = xs;
    v[L - 1] = xs;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = L - 2; l >= 0; --l) {
      x = (g.rp[l] - g.f[l] * xp) - g.d[l] * x;
      v[l] = x;
    }
  }
};

// Picks a segment length for the register kernel; false if this nz is not served.
inline bool thomas_reg_pick(int nz, bool periodic, int* Lout) {
  const int cand[4] = {16, 8, 4, 2};
  for (int q = 0; q < 4; ++q) {
    const int L = cand[q];
    if (nz % L) continue;
    const int S = nz / L;
    if (S < 2 || S > 64) continue;
    if (periodic && (S & (S - 1))) continue;               // cyclic PCR needs a power-of-two number of separators
    *Lout = L;
    return true;
  }
  return false;
}

}  // namespace fb

#if defined(__CUDACC__)
namespace fb {

// store policy: see fft_reg.cuh (FB_STREAM_Y); the z tiles are 64-byte pieces 8*ncol bytes apart
#ifndef FB_STREAM_Z
#define FB_STREAM_Z 1
#endif
__device__ __forceinline__ void st_z(double* p, double v) { if (FB_STREAM_Z) __stcs(p, v); else *p = v; }
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int NLEFT> __device__ __forceinline__ void cp_async_wait_but() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

// ---- thread-block clusters: CL CTAs (one or two per SM) share one tile ------------------------------------------
// A 16-column tile (128-byte row pieces: the access-pattern ceiling at 1024^3 is 2.6 TB/s for 8-column tiles and
// 5.1 TB/s for 16, tools/pattern_bench.cu) of nz = 1024 levels does not fit the registers of one SM.  With CL = 2 the
// CTAs of a cluster split the LEVELS: CTA r keeps segments [r S/CL, (r+1) S/CL) in registers, sends its seven reduced-
// system inputs per segment (R0,F0,G0,RD,FD,DD and the separator's right-hand side) to the peer's shared memory
// (st.shared::cluster), and after ONE cluster barrier per tile both CTAs solve the whole reduced system redundantly
// (2 rows per thread) -- no second exchange, nothing else crosses the SM boundary.
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_store(double* local, unsigned peer, double v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(peer));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

// Distributed z solve (capi.cu, solver_slab_dz): a rank holds n3l consecutive levels of EVERY column.  Pass 1 solves the
// rank-local block T_g y = b (this kernel, CORR = false, coefficients truncated at the slab ends); after the 2G x 2G
// interface system has delivered the neighbours' boundary unknowns x_prev, x_next per column, pass 2 (CORR = true) forms
//     x = y + T_g^{-1} ( -a_first x_prev e_first - c_last x_next e_last )
// in the same sweep: the tile is read as y, the right-hand side is synthesised in registers, and y + correction is stored.
struct ThomasCorr {
  const double* xprev;          // [ncol] last unknown of the rank below (unused where ca == 0)
  const double* xnext;          // [ncol] first unknown of the rank above
  double ca, cc;                // the true couplings a(first level), c(last level) that the local block leaves out
};

// shared memory (doubles) of the kernel below
template <int L, int TI>
inline size_t thomas_reg_smem_doubles(int nz, int maxt, int nbuf, int cl, bool uni) {
  const size_t st = (size_t)(nz / L) * TI;
  const size_t exch = (cl > 1) ? 2 * 7 * st + 3 * st : 13 * st;     // CL > 1: ex[2 parities][7] | pcrA (pcrB, X alias the dead parity)
  return (size_t)nbuf * L * maxt + exch + (uni ? 0 : 3 * (size_t)ThomasReg<L, TI>::tile_rows(nz));
}

// MAXT: upper bound of the block size TI*S/CL (256 -> two blocks per SM, 512 -> one)
// NBUF: fetch depth.  1: the next tile is fetched after phase 1 of the current one; 2: two private slot sets, the tile
// after next is requested as soon as the current one sits in registers.  Measured (B200, v7): NBUF = 2 is SLOWER
// (512^3: 0.59 -> 0.76 ms, 1024^3: 6.36 -> 6.69 ms) -- more requests in flight do not help a kernel that sits at its
// access-pattern ceiling; kept as a compile-time option, default 1.
// CORR (distributed z solve, pass 2; CL = 1, NBUF = 1 only): the tile is NOT fetched -- the right-hand side is the two
// boundary couplings, synthesised in registers -- and y is re-read (L2) and added at the store.
template <int L, int TI, int MAXT, bool UNI, int MINB, int NBUF, int CL, bool CORR = false>
__global__ void __launch_bounds__(MAXT, MINB)
thomas_reg_kernel(long ncol, long ntiles, ThomasArgs T, const double* __restrict__ lam, const double* W,
                  ColGeom og, ThomasCorr corr) {
  using TR = ThomasReg<L, TI>;
  extern __shared__ double smem[];
  const int nz = T.nz, S = T.S;                           // S = segments of a whole column (all CTAs of the cluster)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int st = S * TI;
  const int S_loc = S / CL;
  const unsigned rank = (CL > 1) ? cluster_ctarank() : 0u;
  double* slots = smem;                                   // [NBUF][L][MAXT], private per thread
  double* exbase = slots + (size_t)NBUF * L * MAXT;       // CL = 1: ex[6] ; CL > 1: ex[2][7]
  double* pcrA = exbase + (CL > 1 ? 14 : 6) * (size_t)st; // 3 arrays
  double* pcrB1 = pcrA + 3 * (size_t)st;                  // CL = 1 only: 3 arrays + X
  double* coef = pcrA + (CL > 1 ? 3 : 7) * (size_t)st;    // az | bz | cz at padded rows
  const int lane = tid % TI, s_loc = tid / TI;
  const int s = (int)rank * S_loc + s_loc;                // this thread's segment of the column
  if (!UNI) {
    const int tr = TR::tile_rows(nz);
    for (int k = tid; k < nz; k += nthr) {
      const int r = TR::prow(k);
      coef[r] = __ldg(T.az + k); coef[tr + r] = __ldg(T.bz + k); coef[2 * tr + r] = __ldg(T.cz + k);
    }
    T.az = coef; T.bz = coef + tr; T.cz = coef + 2 * tr; T.padded = 1;
  }
  using CF = typename std::conditional<UNI, CoefUniform<L>, CoefTable<L>>::type;
  // where this thread's levels go: chunk q of the output geometry (one GPU: the work array itself)
  const int k0 = s * L;
  const bool one_chunk = (og.n3l % L) == 0;
  double* obase = nullptr;
  if (one_chunk) { const int q = k0 / og.n3l; obase = og.ptr[q] + og.koff + ncol * (long)(k0 - q * og.n3l); }
  const long tstride = gridDim.x / CL;                    // clusters in the grid

  auto fetch = [&](long tile, int buf) {                   // always commits a group (possibly empty): uniform counting
    if (!CORR && tile < ntiles) {
      const long col = min(tile * TI + lane, ncol - 1);
      const double* src = W + col + (long)k0 * ncol;
      double* sl = slots + (size_t)buf * L * MAXT + tid;
#pragma unroll
      for (int l = 0; l < L; ++l) cp_async8(sl + l * MAXT, src + (long)l * ncol);
    }
    cp_async_commit();
  };

  auto lam_of = [&](long tile) {                           // dead lanes of a ragged last tile: any regular column
    const long col = tile * TI + lane;
    return (col < ncol) ? __ldg(lam + col) : -1.0;
  };
  long tile = blockIdx.x / CL;
  double lm_next = 0.0;
  if (tile < ntiles) { fetch(tile, 0); lm_next = lam_of(tile); }
  if (NBUF == 2) fetch(tile + tstride, 1);
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // coefficients staged; the peer CTA is resident (DSMEM stores below)

  for (int it = 0; tile < ntiles; tile += tstride, ++it) {
    const long col = tile * TI + lane;
    const bool live = col < ncol;
    const double lm = lm_next;                            // loaded one tile ahead
    if (tile + tstride < ntiles) lm_next = lam_of(tile + tstride);
    const bool pin = T.singular && live && (lm == 0.0);
    double v[L];
    const int buf = (NBUF == 2) ? (it & 1) : 0;
    if (NBUF == 2) cp_async_wait_but<1>(); else cp_async_wait_all();
    {
      const double* sl = slots + (size_t)buf * L * MAXT + tid;
#pragma unroll
      for (int l = 0; l < L; ++l) v[l] = CORR ? 0.0 : sl[l * MAXT];
    }
    if (CORR && live) {
      if (s == 0 && corr.ca != 0.0) v[0] = -corr.ca * __ldg(corr.xprev + col);
      if (s == S - 1 && corr.cc != 0.0) v[L - 1] = v[L - 1] - corr.cc * __ldg(corr.xnext + col);
    }
    if (NBUF == 2) fetch(tile + 2 * tstride, buf);        // this slot set is free again (slots are private per thread)

    double* ex = exbase + ((CL > 1) ? (size_t)(it & 1) * 7 * st : 0);
    double* pcrB = (CL > 1) ? ex : pcrB1;                  // CL > 1: the current parity's ex is dead once the rows are built
    double* X = pcrB + 3 * (size_t)st;
    const CF cf(T, s);
    SegRegs<L> g;
    TR::phase1(v, T, cf, lm, lane, s, g, ex);
    // NBUF = 1: the slots have been free since v[] was read, but issuing the copies BEFORE phase 1 keeps 16 address pairs
    // alive across it: the coefficient-table variant then spills (56 bytes) and runs 20 % slower (B200, 512^3, measured)
    if (NBUF == 1) fetch(tile + tstride, 0);
    if (CL > 1) {
      const int o = s * TI + lane;
      ex[6 * st + o] = v[L - 1];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int pr = 1; pr < CL; ++pr) dsmem_store(ex + q * st + o, (rank + pr) % CL, ex[q * st + o]);
      }
      cluster_sync_all();
#pragma unroll
      for (int rr = 0; rr < CL; ++rr) {
        const int s2 = s_loc + rr * S_loc;
        const CF cf2(T, s2);
        TR::reduced_row(ex[6 * st + s2 * TI + lane], ex, pcrA, T, cf2, lm, lane, s2, pin);
      }
    } else {
      __syncthreads();
      TR::reduced_row(v[L - 1], ex, pcrA, T, cf, lm, lane, s, pin);
    }
    __syncthreads();
    double* src = pcrA;
    double* dst = pcrB;
    const int hmax = T.periodic ? S / 2 : S;
    for (int h = 1; h < hmax; h *= 2) {
      bool coupled = false;
#pragma unroll
      for (int rr = 0; rr < CL; ++rr) coupled = TR::pcr_step(src, dst, T, lane, s_loc + rr * S_loc, h) || coupled;
      const int any = __syncthreads_or(coupled ? 1 : 0);
      double* t = src; src = dst; dst = t;
      if (!any) break;                                    // every column of the tile is decoupled already
    }
    // X must not overlay the rows pcr_finish still reads: with CL > 1 it sits in ex[3], pcrB = ex[0..2], pcrA separate
#pragma unroll
    for (int rr = 0; rr < CL; ++rr) TR::pcr_finish(src, X, T, lane, s_loc + rr * S_loc);
    __syncthreads();
    TR::phase3(v, X, T, lane, s, g);
    if (CORR && live) {                                    // x = y + correction; y still sits where x goes (in place)
      const double* yp = W + col + (long)k0 * ncol;
#pragma unroll
      for (int l = 0; l < L; ++l) v[l] += yp[(long)l * ncol];
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
    // CL = 1: X / ex / pcr buffers are rewritten only after the next iteration's barriers.  CL > 1: the next tile uses
    // the other ex parity; the peer writes THIS parity again only after the next cluster barrier, which this CTA
    // reaches after it is done with the tile.
  }
  if (CL > 1) cluster_sync_all();                         // no CTA exits while its peer may still address its shared memory
}

struct ThomasCfgKey { int nz, uni, periodic; };

template <int L, int TI, int MAXT, bool UNI, int MINB, int NBUF, int CL, bool CORR = false>
inline cudaError_t thomas_reg_launch1(long ncol, const ThomasArgs& T, const double* lam, const double* W, const ColGeom& og,
                                      int nsm, cudaStream_t st, const ThomasCorr* corr = nullptr) {
  auto kern = thomas_reg_kernel<L, TI, MAXT, UNI, MINB, NBUF, CL, CORR>;
  const ThomasCorr cr = corr ? *corr : ThomasCorr{nullptr, nullptr, 0.0, 0.0};
  const size_t smem = thomas_reg_smem_doubles<L, TI>(T.nz, MAXT, NBUF, CL, UNI) * sizeof(double);
  const long ntiles = (ncol + TI - 1) / TI;
  const int threads = TI * T.S / CL;
  static int nclusters = 0, cfg_nz = 0;                   // configured once per (kernel, nz)
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  if (nclusters == 0 || cfg_nz != T.nz) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    if (CL > 1) {
      cfg.gridDim = dim3(CL * nsm);
      e = cudaOccupancyMaxActiveClusters(&q, kern, &cfg);
      if (e != cudaSuccess) return e;
    } else {
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem);
      if (e != cudaSuccess) return e;
      q *= nsm;
    }
    if (q < 1) return cudaErrorLaunchOutOfResources;
    nclusters = q; cfg_nz = T.nz;
  }
  const long ncl = ntiles < (long)nclusters ? ntiles : (long)nclusters;
  cfg.gridDim = dim3((unsigned)(ncl * CL));
  return cudaLaunchKernelEx(&cfg, kern, ncol, ntiles, T, lam, W, og, cr);
}

// Tile shapes.  cfg 0: 8 columns, one CTA per tile (v4-v7).  cfg 1: 16 columns, one 512-thread CTA (nz <= 512 at L = 16).
// cfg 2: 16 columns, a cluster of two CTAs splits the levels (nz = 1024: 2 x 512 threads on two SMs; nz = 512: 2 x 256).
template <int L>
inline cudaError_t thomas_reg_dispatch(int cfgsel, long ncol, const ThomasArgs& T, const double* lam, const double* W,
                                       const ColGeom& og, int nsm, cudaStream_t st, bool* served) {
  const int S = T.S;
  *served = true;
#define FB_TL(TI, MAXT, MINB, CL)                                                                                     \
  (T.uniform ? thomas_reg_launch1<L, TI, MAXT, true, MINB, 1, CL>(ncol, T, lam, W, og, nsm, st)                         \
             : thomas_reg_launch1<L, TI, MAXT, false, MINB, 1, CL>(ncol, T, lam, W, og, nsm, st))
  static const bool nbuf2 = [] { const char* e = getenv("FLUTAS_B200_THOMAS_NBUF"); return e && atoi(e) == 2; }();
  if (nbuf2 && cfgsel == 1 && 16 * S <= 512 && (ncol % 16) == 0 &&
      thomas_reg_smem_doubles<L, 16>(T.nz, 512, 2, 1, T.uniform) * 8 <= 227 * 1024)
    return T.uniform ? thomas_reg_launch1<L, 16, 512, true, 1, 2, 1>(ncol, T, lam, W, og, nsm, st)
                     : thomas_reg_launch1<L, 16, 512, false, 1, 2, 1>(ncol, T, lam, W, og, nsm, st);
  const size_t lim = 227 * 1024;
  if (cfgsel == 2 && S % 2 == 0 && (ncol % 16) == 0) {
    const int thr = 16 * S / 2;
    if (thr <= 256 && 2 * (thomas_reg_smem_doubles<L, 16>(T.nz, 256, 1, 2, T.uniform) * 8 + 1024) <= lim + 1024) return FB_TL(16, 256, 2, 2);
    if (thr <= 512 && thomas_reg_smem_doubles<L, 16>(T.nz, 512, 1, 2, T.uniform) * 8 <= lim) return FB_TL(16, 512, 1, 2);
  }
  if (cfgsel == 1 && 16 * S <= 512 && (ncol % 16) == 0) return FB_TL(16, 512, 1, 1);
  if (8 * S <= 256) return FB_TL(8, 256, 2, 1);
  if (8 * S <= 512) return FB_TL(8, 512, 1, 1);
#undef FB_TL
  *served = false;
  return cudaSuccess;
}

// *done = false if this nz is not served (caller falls back to thomas_tile / the generic kernels).
inline int thomas_reg_run(long ncol, int nz, const double* az, const double* bz, const double* cz, const double* lam,
                          const double* W, double* Wout, const ColGeom* out, bool periodic, int singular, int nsm,
                          const ThomasArgs* uni, cudaStream_t st, bool* done) {
  *done = false;
  int L = 0;
  if (!thomas_reg_pick(nz, periodic, &L)) return 0;
  ThomasArgs T;
  T.nz = nz; T.S = nz / L; T.periodic = periodic ? 1 : 0; T.singular = singular; T.az = az; T.bz = bz; T.cz = cz;
  T.padded = 0;
  T.uniform = 0;
  if (uni && uni->uniform) {
    T.uniform = 1; T.a0 = uni->a0; T.b0 = uni->b0; T.a_first = uni->a_first; T.b_first = uni->b_first;
    T.b_last = uni->b_last; T.c_last = uni->c_last;
  }
  ColGeom og;
  if (out) og = *out;
  else { for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = Wout; og.n3l = nz; og.koff = 0; }
  // FLUTAS_B200_THOMAS_CFG = 0 / 1 / 2 forces a tile shape (see thomas_reg_dispatch); default: two-CTA clusters with
  // 16-column tiles when a column needs more than 512 threads at 8 columns... (set from measurements)
  static const int cfg_env = [] { const char* e = getenv("FLUTAS_B200_THOMAS_CFG"); return e ? atoi(e) : -1; }();
  // default: 16-column tiles in one 512-thread CTA whenever a column fits (S <= 32 segments); measured on B200 against
  // 8-column tiles: 512^3 0.585 -> 0.494 ms, 1024x512x512 1.30 -> 0.97 ms, 1024x1024x512 2.99 -> 1.80 ms.  The cluster
  // variant (cfg 2) is correct but slower (1024^3: 8.1 vs 6.4 ms) and stays an option only.
  int cfgsel = cfg_env >= 0 ? cfg_env : ((16 * T.S <= 512) ? 1 : 0);
  bool served = false;
  cudaError_t e = cudaSuccess;
  switch (L) {
    case 2: e = thomas_reg_dispatch<2>(cfgsel, ncol, T, lam, W, og, nsm, st, &served); break;
    case 4: e = thomas_reg_dispatch<4>(cfgsel, ncol, T, lam, W, og, nsm, st, &served); break;
    case 8: e = thomas_reg_dispatch<8>(cfgsel, ncol, T, lam, W, og, nsm, st, &served); break;
    default: e = thomas_reg_dispatch<16>(cfgsel, ncol, T, lam, W, og, nsm, st, &served); break;
  }
  if (e != cudaSuccess) return (int)e;
  *done = served;
  return 0;
}

// Rank-local block of the distributed z solve on a GENERAL z grid (coefficient tables az, bz, cz of the nz = n3l local rows,
// couplings out of the block already zeroed): L = 16, 16-column tiles, S = nz/16 in 2..32; block size fitted to S so that
// several CTAs share an SM.  corr = nullptr: pass 1; else pass 2.  *done = false: shape not served.
inline bool thomas_reg_local_ok(int nz, long ncol) { return nz % 16 == 0 && nz / 16 >= 2 && nz / 16 <= 32 && (ncol % 16) == 0; }
inline int thomas_reg_local_run(long ncol, int nz, const double* az, const double* bz, const double* cz, const double* lam, double* W,
                                int singular, int nsm, const ThomasCorr* corr, cudaStream_t st, bool* done) {
  *done = false;
  if (!thomas_reg_local_ok(nz, ncol)) return 0;
  ThomasArgs T;
  T.nz = nz; T.S = nz / 16; T.periodic = 0; T.singular = singular; T.az = az; T.bz = bz; T.cz = cz; T.padded = 0; T.uniform = 0;
  ColGeom og;
  for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = W;
  og.n3l = nz; og.koff = 0;
  cudaError_t e;
  if (T.S <= 8) e = corr ? thomas_reg_launch1<16, 16, 128, false, 4, 1, 1, true>(ncol, T, lam, W, og, nsm, st, corr)
                         : thomas_reg_launch1<16, 16, 128, false, 4, 1, 1, false>(ncol, T, lam, W, og, nsm, st);
  else if (T.S <= 16) e = corr ? thomas_reg_launch1<16, 16, 256, false, 2, 1, 1, true>(ncol, T, lam, W, og, nsm, st, corr)
                               : thomas_reg_launch1<16, 16, 256, false, 2, 1, 1, false>(ncol, T, lam, W, og, nsm, st);
  else e = corr ? thomas_reg_launch1<16, 16, 512, false, 1, 1, 1, true>(ncol, T, lam, W, og, nsm, st, corr)
                : thomas_reg_launch1<16, 16, 512, false, 1, 1, 1, false>(ncol, T, lam, W, og, nsm, st);
  if (e != cudaSuccess) return (int)e;
  *done = true;
  return 0;
}

}  // namespace fb
#endif
