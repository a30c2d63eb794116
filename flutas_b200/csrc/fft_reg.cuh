// fft_reg.cuh -- kernels around reg_fft.cuh: x lines (contiguous) and y lines (stride n1), persistent blocks.
//
//   xfft_reg_kernel : T = N/32 threads per line.  T <= 32: a line lives inside one warp, so every exchange is
//                     followed by __syncwarp() only -- warps of a block never wait for each other and their
//                     load / compute / store phases overlap.  Physical side: 8-byte accesses (the halo'd p of
//                     the caller is only 8-byte aligned); spectral side: 16-byte accesses, (Re,Im) interleaved.
//   yfft_reg_kernel : lanes = TB consecutive i, so the strided y lines are read and written coalesced; the T
//                     threads of a line sit in different warps -> block barriers.  The spectral side goes
//                     through SpecGeom: on several GPUs its stores ARE the exchange (geom.cuh).
//
// Reference counterparts: fft() forward/backward on the x and y pencils (src/solver_cpu.f90:59,65,86,89) and
// the pack/unpack loops of the transposes around them.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "geom.cuh"
#include "reg_fft.cuh"

namespace fb {

// Tables staged in shared memory by every block: the pass twiddles (tw1, tw2), the split/merge twiddles wN[M] and,
// for Makhoul (NN/DD) lines of length <= 1024, wQ[M+1].  All are indexed j + T u or k + (t-1) Ns, i.e. one run-time
// base plus immediates (the v7 kernels fetched wN/wQ from global memory: 62 of their 94 LDG per thread and tile).
template <class S, bool MK>
struct RegTw {
  static constexpr int M = S::M;
  static constexpr int n1 = (S::NP > 1) ? (S::radix(1) - 1) * S::ns(1) : 0;
  static constexpr int n2 = (S::NP > 2) ? (S::radix(2) - 1) * S::ns(2) : 0;
  static constexpr int nN = M;
  static constexpr bool QSM = MK && (M <= 512);
  static constexpr int nQ = QSM ? M + 1 : 0;
  static constexpr int total = n1 + n2 + nN + nQ + ((n1 + n2 + nQ) & 1);     // cpx entries (16 bytes each)
};

// exchange buffer of one x line: M slots of (re,im); XOR-swizzled (reg_fft.cuh: rf_swz) or one pad slot per 16.
// Measured on B200 (profiles/r02_swz_ab.log): the swizzle wins where a line spans two warps (N = 2048: x fwd 2.41 -> 2.31,
// x inv 3.10 -> 2.87 ms) and loses in the warp-per-line kernels (N = 1024: 3.28 -> 3.38 ms; their 128-register budget has
// no room for the extra address registers, and keeping the compiler from hoisting them out of the line loop only gets back
// to parity) -> chosen per length.  FB_XBUF_PAD=1 / FB_XBUF_SWZ=1 force one (A/B builds).
// FB_XFETCH_AHEAD=1 (A/B builds): the forward x kernel issues the next line's loads into the registers the head passes have
// freed, under the pair pass of the current line.  Costs 200-330 bytes of spills per thread at the 128-register budget.
#ifndef FB_XFETCH_AHEAD
#define FB_XFETCH_AHEAD 0
#endif
#ifndef FB_PAIR_SPLIT
#define FB_PAIR_SPLIT 1
#endif
#ifndef FB_PAIR_Y
#define FB_PAIR_Y 1
#endif
#ifndef FB_PAIR_YB
#define FB_PAIR_YB 1
#endif
#ifndef FB_PAIR_MERGE
#define FB_PAIR_MERGE 1
#endif
#ifndef FB_XBUF_PAD
#define FB_XBUF_PAD 0
#endif
#ifndef FB_XBUF_SWZ
#define FB_XBUF_SWZ 0
#endif
template <bool XOR_>
struct XLineBufT {
  static constexpr bool XOR = XOR_;
  double2* b;
  static FB_CX int off(int c) { return XOR ? rf_swzoff(c) : rf_padoff(c); }
  static FB_CX int slot(int m) { return XOR ? rf_swz(m) : rf_pad(m); }
  static FB_CX int length(int M) { return XOR ? M : M + M / 16; }
  static FB_CX int at(int bs, int coff) { return XOR ? (bs ^ coff) : (bs + coff); }
  static FB_CX int offsub(int a, int b) { return XOR ? (rf_swz(a) ^ rf_swz(b)) : (rf_padoff(a) - rf_padoff(b)); }   // offset code of +a - b
  __device__ __forceinline__ int addr(int bs, int coff) const { return at(bs, coff); }                               // ld/st(addr, 0)
  __device__ __forceinline__ int base(int pos) const { return slot(pos); }
  __device__ __forceinline__ void st(int bs, int coff, double r, double i) const { b[at(bs, coff)] = make_double2(r, i); }
  __device__ __forceinline__ void ld(int bs, int coff, double& r, double& i) const {
    const double2 v = b[at(bs, coff)];
    r = v.x; i = v.y;
  }
};
template <int N>
using XLineBuf = XLineBufT<(FB_XBUF_SWZ || N >= 2048) && !FB_XBUF_PAD>;

// exchange buffer of a y tile: [slot][lane]
template <int TB>
struct YTileBuf {
  static constexpr bool XOR = false;
  double2* b;   // already offset by the lane
  static FB_CX int off(int c) { return rf_padoff(c); }
  static FB_CX int offsub(int a, int b) { return rf_padoff(a) - rf_padoff(b); }
  __device__ __forceinline__ int addr(int bs, int coff) const { return bs + coff * TB; }
  __device__ __forceinline__ int base(int pos) const { return rf_pad(pos) * TB; }
  __device__ __forceinline__ void st(int bs, int coff, double r, double i) const { b[bs + coff * TB] = make_double2(r, i); }
  __device__ __forceinline__ void ld(int bs, int coff, double& r, double& i) const {
    const double2 v = b[bs + coff * TB];
    r = v.x; i = v.y;
  }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Store policy of the strided sides (compile-time, for A/B builds).  tools/pattern_bench.cu (same tiles, no arithmetic)
// puts the ceiling of plain stores above st.cs for 64-byte pieces (halves of a 128-byte line merge in L2 instead of
// leaving it evict-first), but the transform kernels are not at that ceiling: measured on B200, 512^3 and 1024^3,
// st.cs is 0-2 % faster than plain stores on every stage -> st.cs stays the default.
// y lines of N = 1024 with 8 values per thread (YRegShape, RR = 8: 64 registers, 32 warps per SM, schedule 8 x 8 x 8) were the
// default until the 16-value schedule {16,16,2} got the pair passes (two exchanges and four barriers per tile instead of
// three and seven): measured on B200 at 1024^3 (profiles/r02_ypair_ab.log) y fwd 4.34 -> 4.00 ms, y inv 4.48 -> 3.98 ms.
// FLUTAS_B200_Y8=1 selects the 8-value kernels.
#ifndef FB_Y8_DEFAULT
#define FB_Y8_DEFAULT 0
#endif
#ifndef FB_X8_DEFAULT
#define FB_X8_DEFAULT 0
#endif
#ifndef FB_STREAM_Y
#define FB_STREAM_Y 1
#endif
#ifndef FB_STREAM_X
#define FB_STREAM_X 1
#endif
#ifndef FB_LDCS_Y
#define FB_LDCS_Y 0
#endif
__device__ __forceinline__ void st_y(double* p, double v) { if (FB_STREAM_Y) __stcs(p, v); else *p = v; }
__device__ __forceinline__ double ld_y(const double* p) { return FB_LDCS_Y ? __ldcs(p) : *p; }
__device__ __forceinline__ void st_x2(double2* p, double2 v) { if (FB_STREAM_X) __stcs(p, v); else *p = v; }


// copies the tables to shared memory at `sm` in the order tw1 | tw2 | wN | wQ (offsets: RegTw).  The kernels form the
// table pointers as plain local variables: handing them around inside a struct made nvcc 12.9 lose the shared
// address space of one member (a generic load from the bare shared offset -> illegal address).
template <class S, bool MK>
__device__ __forceinline__ void reg_stage_tw(cpx* sm, const RegPlan& P, int tid, int nthr) {
  using TW = RegTw<S, MK>;
  auto copy = [&](cpx* d, const cpx* g, int n) {
    const double2* g2 = reinterpret_cast<const double2*>(g);
    for (int q = tid; q < n; q += nthr) reinterpret_cast<double2*>(d)[q] = __ldg(g2 + q);
  };
  copy(sm, S::R == 16 ? P.tw[1] : P.tw8[1], TW::n1);
  copy(sm + TW::n1, S::R == 16 ? P.tw[2] : P.tw8[2], TW::n2);
  copy(sm + TW::n1 + TW::n2, P.wN, TW::nN);
  if (TW::QSM) copy(sm + TW::n1 + TW::n2 + TW::nN, P.wQ, TW::nQ);
}
#define FB_REG_TABLES(S, MK, smbase, P)                                                       \
  const cpx* s_tw1 = reinterpret_cast<const cpx*>(smbase);                                     \
  const cpx* s_tw2 = s_tw1 + RegTw<S, MK>::n1;                                                 \
  const cpx* s_wN = s_tw2 + RegTw<S, MK>::n2;                                                  \
  const cpx* s_wQ = RegTw<S, MK>::QSM ? (s_wN + RegTw<S, MK>::nN) : P.wQ;                      \
  const cpx* tw[RF_MAXPASS] = {nullptr, s_tw1, s_tw2};

template <int N, bool MK, int RR = 16>
constexpr size_t xfft_reg_smem() {
  constexpr int M = N / 2;
  return (size_t)RegTw<RegSched<M, RR>, MK>::total * sizeof(cpx) + (size_t)(256 / RegSched<M, RR>::T) * XLineBuf<N>::length(M) * sizeof(double2);
}

// KC: 0 = periodic (R2HC / HC2R), 1 = Makhoul (NN / DD), 2 = types IV (ND / DN)
// RR = 8 (N = 1024 only, A/B variant FLUTAS_B200_X8): 64 threads = two warps per line at 8 values each, <= 64 registers,
// four 256-thread blocks = 32 warps per SM instead of 16.
template <int N, bool FWD, int KC, int RR = 16>
__global__ void __launch_bounds__(256, RR == 8 ? 4 : 2)
xfft_reg_kernel(RegPlan P, const double* __restrict__ src, LineGeom gs, double* __restrict__ dst, LineGeom gd, double scale) {
  constexpr bool MK = (KC == 1), IV = (KC == 2);
  constexpr int M = N / 2;
  using S = RegSched<M, RR>;
  constexpr int T = S::T, R = S::R;
  constexpr bool WARP = (T <= 32);
  constexpr bool PAIR = FWD && !IV && FB_PAIR_SPLIT && reg_has_pair_pass<S>();
  // backward: measured on B200 (profiles/r02_pairb_ab.log) the transposed schedule wins on DCT/DST lines (N = 1024: 2.27 -> 1.96 ms,
  // N = 2048: 2.86 -> 2.40 ms) and loses on periodic ones (N = 1024: 3.21 -> 3.48 ms) -> Makhoul kinds only
  constexpr bool PAIRB = !FWD && !IV && (MK || FB_PAIR_MERGE > 1) && FB_PAIR_MERGE && reg_has_pair_pass<S>();
  constexpr int GT = WARP ? 32 : 256;                 // threads that synchronise with each other
  constexpr int LG = GT / T;                          // lines per group
  constexpr int BUFL = XLineBuf<N>::length(M);
  extern __shared__ double2 smem2[];
  double2* bufs = smem2 + RegTw<S, MK>::total;
  const int tid = threadIdx.x;
  reg_stage_tw<S, MK>(reinterpret_cast<cpx*>(smem2), P, tid, 256);
  __syncthreads();
  FB_REG_TABLES(S, MK, smem2, P)
  const int gi = WARP ? (tid >> 5) : 0, tg = tid % GT;
  const int lw = tg / T, j = tg % T;
  const XLineBuf<N> xb{bufs + (size_t)(gi * LG + lw) * BUFL};
  // a line of N = 2048 needs T = 64 threads = two whole warps: they meet at their own named barrier, so the four lines of
  // a block never wait for each other (B200, 2048 x 2048 x 128: x fwd 2.92 -> 2.48 ms, x inv 3.35 -> 3.13 ms against __syncthreads)
  auto sync = [lw] {
    if (WARP) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + lw), "n"(T) : "memory");
  };
  const int kind = (MK || IV) ? P.kind : (int)KIND_PP;
  const bool dn = IV && (P.kind == KIND_DN);
  const long nlines = gs.nlines;
  const long ngroups = (nlines + LG - 1) / LG;
  const long gstride = WARP ? (long)gridDim.x * 8 : (long)gridDim.x;
  // Periodic lines whose SECOND element is 16-byte aligned (the halo'd p of the caller: interior starts at an odd
  // element) are packed one element late, z_m = x_{2m+1} + i x_{2m+2} (indices mod N), so that all but one of
  // the physical accesses are aligned 16-byte vectors.  A cyclic shift only multiplies mode k by a unit
  // phase, which commutes with the (real, per-mode) z solve; the backward kernel stores with the same shift.
  const LineGeom& gph = FWD ? gs : gd;
  const bool even_strides = (gph.sj % 2 == 0) && (gph.sk % 2 == 0);
  const long el0 = (long)(reinterpret_cast<uintptr_t>(FWD ? (const void*)src : (const void*)dst) / 8) + gph.off0;
  const bool al1 = even_strides && ((el0 + 1) % 2 == 0);       // element 1 of every line is 16-byte aligned
  const bool al0 = even_strides && (el0 % 2 == 0);             // element 0 is
  const bool shift = !MK && !IV && al1;
  // DCT/DST lines: the Makhoul permutation makes per-thread accesses 32 bytes apart, so the physical side is
  // read / written in natural order as aligned 16-byte pairs and permuted through the line's exchange buffer.
  const bool viabuf = MK && (al0 || al1);
  // natural-order pair q of a line: elements (ea, eb); `vec` = one aligned 16-byte access at element ea
  auto pair_elems = [&](int q, int& ea, int& eb, bool& vec) {
    if (al0) { ea = 2 * q; eb = 2 * q + 1; vec = true; }
    else if (q == M - 1) { ea = N - 1; eb = 0; vec = false; }
    else { ea = 2 * q + 1; eb = 2 * q + 2; vec = true; }
  };
  auto slot_sign = [&](int e, double& sgn) {                   // slot (doubles) of physical element e in the buffer
    int m, part;
    elem_to_slot(kind, N, e, m, part, sgn);
    return 2 * XLineBuf<N>::slot(m) + part;
  };
  double* lbuf = reinterpret_cast<double*>(xb.b);
  // forward: the raw loads of a line (fetch) are separate from their arrangement (signs / Makhoul permutation), so that the
  // pair-pass kernels can issue the NEXT line's loads as soon as the registers are free (after the head passes have stored to
  // the exchange buffer) and let them fly under the pair pass, the split and the stores of the current line
  auto fetch = [&](long lcn, double (&re)[R], double (&im)[R]) {
    const double* ps = src + line_offset(gs, lcn);
    if (shift) {
      const double2* pa = reinterpret_cast<const double2*>(ps + 1);
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int m = j + T * u;
        if (u == R - 1 && j == T - 1) { re[u] = ps[N - 1]; im[u] = ps[0]; }
        else { const double2 v = pa[m]; re[u] = v.x; im[u] = v.y; }
      }
    } else if (!MK && !IV) {
#pragma unroll
      for (int u = 0; u < R; ++u) { const int m = j + T * u; re[u] = ps[2 * m]; im[u] = ps[2 * m + 1]; }
    } else if (viabuf) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        int ea, eb; bool vec;
        pair_elems(j + T * u, ea, eb, vec);
        if (vec) { const double2 v = *reinterpret_cast<const double2*>(ps + ea); re[u] = v.x; im[u] = v.y; }
        else { re[u] = ps[ea]; im[u] = ps[eb]; }
      }
    } else {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        int e0, e1; double s0, s1;
        reg_phys_slots(kind, N, j + T * u, e0, e1, s0, s1);
        re[u] = ps[e0]; im[u] = ps[e1];
      }
    }
  };
  auto arrange = [&](double (&re)[R], double (&im)[R]) {
    if (!MK && !IV) return;
    if (viabuf) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        int ea, eb; bool vec; double sa, sb;
        pair_elems(j + T * u, ea, eb, vec);
        const int pa = slot_sign(ea, sa), pb = slot_sign(eb, sb);
        lbuf[pa] = sa * re[u]; lbuf[pb] = sb * im[u];
      }
      sync();
      reg_gather<S>(re, im, j, xb);
      sync();
    } else {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        int e0, e1; double s0, s1;
        reg_phys_slots(kind, N, j + T * u, e0, e1, s0, s1);
        re[u] = s0 * re[u]; im[u] = s1 * im[u];
      }
    }
  };
  double re[R], im[R];
  bool fetched = false;                                // (re, im) already hold this iteration's raw line
  for (long g = WARP ? (long)blockIdx.x * 8 + gi : (long)blockIdx.x; g < ngroups; g += gstride) {
    const long line = g * LG + lw;
    const bool live = line < nlines;
    const long lc = live ? line : nlines - 1;
    {                                                  // a later group's line -> L2 while this one is transformed
      const long ln = min(line + (PAIR && FB_XFETCH_AHEAD ? 2 : 1) * gstride * LG, nlines - 1);
      const double* pn = src + line_offset(gs, ln) + (N / T) * j;
#pragma unroll
      for (int q = 0; q < (N / T) / 16; ++q) prefetch_l2(pn + 16 * q);
    }
    if (FWD) {
      if (!fetched) fetch(lc, re, im);
      arrange(re, im);
      if (IV) reg_iv_pre<S, false>(re, im, j, P.wQ);
      if constexpr (PAIR) {                             // last pass on symmetric butterfly pairs: split in registers, stored from there
        reg_fft_passes_head<S, -1>(re, im, j, tw, xb, sync);
        sync();
        if (FB_XFETCH_AHEAD) {
          fetched = g + gstride < ngroups;
          if (fetched) fetch(min((g + gstride) * LG + lw, nlines - 1), re, im);
        }
        double2* pd = reinterpret_cast<double2*>(dst + line_offset(gd, lc));
        reg_pair_pass_split<S, MK>(j, tw[S::NP - 1], s_wN, s_wQ, xb, [&](int k, double xr, double xi) {
          if (live) st_x2(pd + k, make_double2(scale * xr, scale * xi));
        });
      } else {
        reg_fft_passes<S, -1>(re, im, j, tw, xb, sync);
        if (IV) {
          reg_iv_post<S, true>(re, im, j, s_wN, dn);
        } else {
          reg_scatter_modes<S>(re, im, j, xb);
          sync();
          reg_split<S, MK>(re, im, j, s_wN, s_wQ, xb);
        }
        if (live) {
          double2* pd = reinterpret_cast<double2*>(dst + line_offset(gd, line));
#pragma unroll
          for (int u = 0; u < R; ++u) st_x2(pd + (j + T * u), make_double2(scale * re[u], scale * im[u]));
        }
      }
    } else {
      const double2* ps = reinterpret_cast<const double2*>(src + line_offset(gs, lc));
      if constexpr (PAIRB) {                            // rows of (k, M - k) merged in registers, transposed schedule
        reg_pair_merge_pass<S, MK>(j, tw[S::NP - 1], s_wN, s_wQ, xb, [&](int k, double& xr, double& xi) {
          const double2 v = ps[k];
          xr = v.x; xi = v.y;
        });
        reg_fft_passes_T_tail<S, +1>(re, im, j, tw, xb, sync);
        if (viabuf) sync();                            // the last pass read the buffer the permutation below rewrites
      } else {
#pragma unroll
        for (int u = 0; u < R; ++u) { const double2 v = ps[j + T * u]; re[u] = dn ? v.y : v.x; im[u] = dn ? v.x : v.y; }
        if (IV) {
          reg_iv_pre<S, true>(re, im, j, P.wQ);
        } else {
          reg_scatter_modes<S>(re, im, j, xb);
          sync();
          reg_merge<S, MK>(re, im, j, s_wN, s_wQ, xb);
          sync();
        }
        reg_fft_passes<S, +1>(re, im, j, tw, xb, sync);
        if (IV) reg_iv_post<S, false>(re, im, j, s_wN, dn);
      }
      if (viabuf) {                                    // packed element m = slot m, then read back in natural order
        reg_scatter_modes<S>(re, im, j, xb);
        sync();
#pragma unroll
        for (int u = 0; u < R; ++u) {
          int ea, eb; bool vec; double sa, sb;
          pair_elems(j + T * u, ea, eb, vec);
          const int pa = slot_sign(ea, sa), pb = slot_sign(eb, sb);
          re[u] = sa * scale * lbuf[pa]; im[u] = sb * scale * lbuf[pb];
        }
      }
      if (live) {
        double* pd = dst + line_offset(gd, line);
        if (viabuf) {
#pragma unroll
          for (int u = 0; u < R; ++u) {
            int ea, eb; bool vec;
            pair_elems(j + T * u, ea, eb, vec);
            if (vec) *reinterpret_cast<double2*>(pd + ea) = make_double2(re[u], im[u]);
            else { pd[ea] = re[u]; pd[eb] = im[u]; }
          }
        } else if (shift) {
          double2* pa = reinterpret_cast<double2*>(pd + 1);
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const int m = j + T * u;
            if (u == R - 1 && j == T - 1) { pd[N - 1] = scale * re[u]; pd[0] = scale * im[u]; }
            else pa[m] = make_double2(scale * re[u], scale * im[u]);
          }
        } else if (!MK && !IV) {
#pragma unroll
          for (int u = 0; u < R; ++u) { const int m = j + T * u; pd[2 * m] = scale * re[u]; pd[2 * m + 1] = scale * im[u]; }
        } else {
#pragma unroll
          for (int u = 0; u < R; ++u) {
            int e0, e1; double s0, s1;
            reg_phys_slots(kind, N, j + T * u, e0, e1, s0, s1);
            pd[e0] = s0 * scale * re[u]; pd[e1] = s1 * scale * im[u];
          }
        }
      }
    }
    sync();                                            // the buffer is rewritten by the next group
  }
}

// ---- y lines ------------------------------------------------------------------------------------
// WIDE: 512-thread blocks when a line needs >= 32 threads (N >= 1024), i.e. 16 lanes = 128-byte row pieces but one
// block per SM; !WIDE: always 256 threads (8 lanes at N = 1024, two blocks per SM).
// RR = 8 (N = 1024 only): 64 threads per line at 8 values each -> 8 lanes in a 512-thread block at <= 64 registers, two
// blocks = 32 warps per SM (twice the warps of the RR = 16 shape with the same shared-memory footprint and pass count).
template <int N, bool WIDE, bool MK, int RR = 16>
struct YRegShape {
  using S = RegSched<N / 2, RR>;
  static constexpr int T = S::T;
  // RR = 8 && WIDE: 16 lanes x 64 threads per line = ONE 1024-thread block per SM, 128-byte row pieces (A/B variant)
  static constexpr int NTMAX = (RR == 8) ? (WIDE ? 1024 : 512) : (WIDE && T >= 32) ? 512 : 256;
  static constexpr int TB = (NTMAX / T > 32) ? 32 : NTMAX / T;
  static constexpr int NT = TB * T;
  static constexpr int MINB = (RR == 8) ? (WIDE ? 1 : 2) : (NT > 256) ? 1 : 2;
  static constexpr size_t smem = (size_t)RegTw<S, MK>::total * sizeof(cpx) + (size_t)(N / 2 + N / 32) * TB * sizeof(double2);
};

// element `c * stride` past p with a compile-time c: one IMAD.WIDE (32-bit stride times immediate plus 64-bit base)
// (stride in BYTES as an unsigned 32-bit value: a signed or element stride costs a high-word fix-up per address)
__device__ __forceinline__ const double* yrow(const double* p, unsigned sbytes, int c) {
  const char* q = reinterpret_cast<const char*>(p);
  return reinterpret_cast<const double*>(c >= 0 ? q + (size_t)sbytes * (unsigned)c : q - (size_t)sbytes * (unsigned)(-c));
}
__device__ __forceinline__ double* yrow(double* p, unsigned sbytes, int c) {
  char* q = reinterpret_cast<char*>(p);
  return reinterpret_cast<double*>(c >= 0 ? q + (size_t)sbytes * (unsigned)c : q - (size_t)sbytes * (unsigned)(-c));
}

template <int N, bool FWD, bool WIDE, int KC, int RR>
__global__ void __launch_bounds__(YRegShape<N, WIDE, KC == 1, RR>::NT, YRegShape<N, WIDE, KC == 1, RR>::MINB)
yfft_reg_kernel(RegPlan P, double* W, int n1, int ntile_i, long ntiles, SpecGeom sg) {
  constexpr bool MK = (KC == 1), IV = (KC == 2);
  constexpr int M = N / 2;
  using S = RegSched<M, RR>;
  using Y = YRegShape<N, WIDE, MK, RR>;
  using MR = MkRows<N, RR>;
  constexpr int T = S::T, R = S::R, TB = Y::TB, NT = Y::NT;
  constexpr bool PAIR = !IV && (FWD ? FB_PAIR_Y : FB_PAIR_YB) && reg_has_pair_pass<S>();
  extern __shared__ double2 smem2[];
  double2* buf = smem2 + RegTw<S, MK>::total;
  const int tid = threadIdx.x;
  reg_stage_tw<S, MK>(reinterpret_cast<cpx*>(smem2), P, tid, NT);
  __syncthreads();
  FB_REG_TABLES(S, MK, smem2, P)
  const int lane = tid % TB, j = tid / TB;
  const YTileBuf<TB> xb{buf + lane};
  auto sync = [] { __syncthreads(); };
  const double sdd = (MK && P.kind == KIND_DD) ? -1.0 : 1.0;   // DST-II/III through the DCT: odd physical elements change sign
  const bool dn = IV && (P.kind == KIND_DN);
  const unsigned sstride = (unsigned)sg.n1l * 8u, pstride = (unsigned)n1 * 8u;      // row strides in bytes (< 4 GB)
  // tile -> (i-tile, k) with 32-bit arithmetic (the tile count of one GPU is far below 2^31): the 64-bit divisions cost
  // a ~100-instruction subroutine per tile and showed up with 3-4 % of the stall samples in the r02 source-level profile
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const unsigned ut = (unsigned)tile, uk = ut / (unsigned)ntile_i;
    const int ti = (int)(ut - uk * (unsigned)ntile_i);
    const long k = (long)uk;
    const int i0 = ti * TB;
    const bool live = (i0 + lane) < n1;
    const int il = i0 + (live ? lane : 0);
    double* base = W + (long)n1 * N * k + il;               // physical side
    double* sbase = spec_base(sg, il, N, k);                 // spectral side (pencil chunk of the exchange)
    {                                                        // next tile -> L2 while this one is transformed
      const long tn = tile + gridDim.x;
      if (tn < ntiles) {
        const unsigned utn = (unsigned)tn, ukn = utn / (unsigned)ntile_i;
        const int tin = (int)(utn - ukn * (unsigned)ntile_i);
        const long kn = (long)ukn;
        const int iln = min(tin * TB, n1 - 1);
        const double* pn = FWD ? (W + (long)n1 * N * kn + iln) : spec_base(sg, iln, N, kn);
        const unsigned sn = FWD ? pstride : sstride;
        constexpr int LP = (TB * 8 > 128) ? TB * 8 / 128 : 1;          // 128-byte lines per row piece
        constexpr int OPS = N * LP / NT;                               // prefetches per thread (N * LP >= NT)
#pragma unroll
        for (int q = 0; q < OPS; ++q) {
          const int op = tid + NT * q;
          prefetch_l2(yrow(pn, sn, op / LP) + 16 * (op % LP));
        }
      }
    }
    double re[R], im[R];
    if (FWD) {
      if (IV) {                                              // z_m = x_{2m} + i x_{N-1-2m}, m = j + T u (DN: the two swapped)
        const double* pe = yrow(base, pstride, 2 * j);
        const double* po = yrow(base, pstride, N - 1 - 2 * j);
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const double a = *yrow(pe, pstride, 2 * T * u), b = *yrow(po, pstride, -2 * T * u);
          re[u] = dn ? b : a; im[u] = dn ? a : b;
        }
        reg_iv_pre<S, false>(re, im, j, P.wQ);
      } else if (!MK) {
        const double* p0 = yrow(base, pstride, 2 * j);
#pragma unroll
        for (int u = 0; u < R; ++u) { re[u] = ld_y(yrow(p0, pstride, 2 * T * u)); im[u] = ld_y(yrow(p0, pstride, 2 * T * u + 1)); }
      } else {
        const double* plo = yrow(base, pstride, MR::base_lo(j));
        const double* phi = yrow(base, pstride, MR::base_hi(j));
#pragma unroll
        for (int u = 0; u < R; ++u) {
          if (!MR::upper(u)) { re[u] = *yrow(plo, pstride, MR::off0(u)); im[u] = *yrow(plo, pstride, MR::off1(u)); }
          else { re[u] = sdd * *yrow(phi, pstride, MR::off0(u)); im[u] = sdd * *yrow(phi, pstride, MR::off1(u)); }
        }
      }
      if constexpr (PAIR) {                               // last pass on symmetric butterfly pairs, split in registers
        reg_fft_passes_head<S, -1>(re, im, j, tw, xb, sync);
        sync();
        reg_pair_pass_split<S, MK>(j, tw[S::NP - 1], s_wN, s_wQ, xb, [&](int k, double xr, double xi) {
          if (live) {
            double* pr = yrow(sbase, sstride, 2 * k);
            st_y(pr, xr);
            st_y(yrow(pr, sstride, 1), xi);
          }
        });
      } else {
        reg_fft_passes<S, -1>(re, im, j, tw, xb, sync);
        if (IV) {
          reg_iv_post<S, true>(re, im, j, s_wN, dn);
        } else {
          reg_scatter_modes<S>(re, im, j, xb);
          sync();
          reg_split<S, MK>(re, im, j, s_wN, s_wQ, xb);
        }
        if (live) {
          double* ps = yrow(sbase, sstride, 2 * j);
#pragma unroll
          for (int u = 0; u < R; ++u) {
            st_y(yrow(ps, sstride, 2 * T * u), re[u]);
            st_y(yrow(ps, sstride, 2 * T * u + 1), im[u]);
          }
        }
      }
    } else {
      if constexpr (PAIR) {                               // rows of (k, M - k) merged in registers, transposed schedule
        reg_pair_merge_pass<S, MK>(j, tw[S::NP - 1], s_wN, s_wQ, xb, [&](int k, double& xr, double& xi) {
          const double* pr = yrow(sbase, sstride, 2 * k);
          xr = ld_y(pr); xi = ld_y(yrow(pr, sstride, 1));
        });
        reg_fft_passes_T_tail<S, +1>(re, im, j, tw, xb, sync);
      } else {
        {
          const double* ps = yrow(sbase, sstride, 2 * j);
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const double a = ld_y(yrow(ps, sstride, 2 * T * u)), b = ld_y(yrow(ps, sstride, 2 * T * u + 1));
            re[u] = dn ? b : a; im[u] = dn ? a : b;
          }
        }
        if (IV) {
          reg_iv_pre<S, true>(re, im, j, P.wQ);
        } else {
          reg_scatter_modes<S>(re, im, j, xb);
          sync();
          reg_merge<S, MK>(re, im, j, s_wN, s_wQ, xb);
          sync();
        }
        reg_fft_passes<S, +1>(re, im, j, tw, xb, sync);
      }
      if (IV) reg_iv_post<S, false>(re, im, j, s_wN, dn);
      if (live) {
        if (IV) {                                            // packed element k -> (x_{2k}, x_{N-1-2k}); DN: (x_{N-1-2k}, x_{2k})
          double* pe = yrow(base, pstride, 2 * j);
          double* po = yrow(base, pstride, N - 1 - 2 * j);
#pragma unroll
          for (int u = 0; u < R; ++u) {
            *yrow(pe, pstride, 2 * T * u) = dn ? im[u] : re[u];
            *yrow(po, pstride, -2 * T * u) = dn ? re[u] : im[u];
          }
        } else if (!MK) {
          double* p0 = yrow(base, pstride, 2 * j);
#pragma unroll
          for (int u = 0; u < R; ++u) { *yrow(p0, pstride, 2 * T * u) = re[u]; *yrow(p0, pstride, 2 * T * u + 1) = im[u]; }
        } else {
          double* plo = yrow(base, pstride, MR::base_lo(j));
          double* phi = yrow(base, pstride, MR::base_hi(j));
#pragma unroll
          for (int u = 0; u < R; ++u) {
            if (!MR::upper(u)) { *yrow(plo, pstride, MR::off0(u)) = re[u]; *yrow(plo, pstride, MR::off1(u)) = im[u]; }
            else { *yrow(phi, pstride, MR::off0(u)) = sdd * re[u]; *yrow(phi, pstride, MR::off1(u)) = sdd * im[u]; }
          }
        }
      }
    }
    sync();
  }
}


template <int N, bool FWD, int KC, int RR = 16>
inline cudaError_t reg_launch_x1(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale,
                                 int nsm, cudaStream_t st) {
  constexpr int M = N / 2, T = RegSched<M, RR>::T;
  constexpr int LPB = 256 / T;                         // lines per block per iteration
  const size_t smem = xfft_reg_smem<N, KC == 1, RR>();
  auto kern = xfft_reg_kernel<N, FWD, KC, RR>;
  static int per_sm = 0;                               // configured once per process (one device per process)
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, 256, smem);
    if (e != cudaSuccess) return e;
    if (q < 1) return cudaErrorLaunchOutOfResources;
    per_sm = q;
  }
  const long nblk = (gs.nlines + LPB - 1) / LPB;
  const long grid = nblk < (long)nsm * per_sm ? nblk : (long)nsm * per_sm;
  kern<<<(unsigned)grid, 256, smem, st>>>(P, src, gs, dst, gd, scale);
  return cudaGetLastError();
}

template <int N, bool FWD>
inline cudaError_t reg_launch_x(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale,
                                int nsm, cudaStream_t st) {
  if constexpr (N == 1024) {                            // 8 values per thread, 32 warps per SM: FLUTAS_B200_X8=1
    static const int x8 = [] { const char* e = getenv("FLUTAS_B200_X8"); return e ? atoi(e) : FB_X8_DEFAULT; }();
    if (x8 && P.tw8[1]) {
      if (kind_is_iv(P.kind)) return reg_launch_x1<N, FWD, 2, 8>(P, src, gs, dst, gd, scale, nsm, st);
      return (P.kind == KIND_PP) ? reg_launch_x1<N, FWD, 0, 8>(P, src, gs, dst, gd, scale, nsm, st)
                                 : reg_launch_x1<N, FWD, 1, 8>(P, src, gs, dst, gd, scale, nsm, st);
    }
  }
  if (kind_is_iv(P.kind)) return reg_launch_x1<N, FWD, 2>(P, src, gs, dst, gd, scale, nsm, st);
  return (P.kind == KIND_PP) ? reg_launch_x1<N, FWD, 0>(P, src, gs, dst, gd, scale, nsm, st)
                             : reg_launch_x1<N, FWD, 1>(P, src, gs, dst, gd, scale, nsm, st);
}

template <int N, bool FWD, bool WIDE, int KC, int RR = 16>
inline cudaError_t reg_launch_y1(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st) {
  using Y = YRegShape<N, WIDE, KC == 1, RR>;
  auto kern = yfft_reg_kernel<N, FWD, WIDE, KC, RR>;
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y::smem);
    if (e != cudaSuccess) return e;
    int q = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, Y::NT, Y::smem);
    if (e != cudaSuccess) return e;
    if (q < 1) return cudaErrorLaunchOutOfResources;
    per_sm = q;
  }
  const int nti = (n1 + Y::TB - 1) / Y::TB;
  const long ntiles = (long)nti * n3;
  const long grid = ntiles < (long)nsm * per_sm ? ntiles : (long)nsm * per_sm;
  kern<<<(unsigned)grid, Y::NT, Y::smem, st>>>(P, W, n1, nti, ntiles, sg);
  return cudaGetLastError();
}

template <int N, bool FWD>
inline cudaError_t reg_launch_y(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, bool wide,
                                cudaStream_t st) {
  const int kc = kind_is_iv(P.kind) ? 2 : (P.kind != KIND_PP) ? 1 : 0;
  if constexpr (N == 1024) {                            // 8 values per thread (see YRegShape); FLUTAS_B200_Y8=0/1 overrides
    static const int y8 = [] { const char* e = getenv("FLUTAS_B200_Y8"); return e ? atoi(e) : FB_Y8_DEFAULT; }();
    static const int y8w = [] { const char* e = getenv("FLUTAS_B200_Y8WIDE"); return e ? atoi(e) : 0; }();
    if (y8 && y8w && P.tw8[1])
      return kc == 2 ? reg_launch_y1<N, FWD, true, 2, 8>(P, W, n1, n3, sg, nsm, st)
           : kc == 1 ? reg_launch_y1<N, FWD, true, 1, 8>(P, W, n1, n3, sg, nsm, st)
                     : reg_launch_y1<N, FWD, true, 0, 8>(P, W, n1, n3, sg, nsm, st);
    if (y8 && !wide && P.tw8[1])
      return kc == 2 ? reg_launch_y1<N, FWD, false, 2, 8>(P, W, n1, n3, sg, nsm, st)
           : kc == 1 ? reg_launch_y1<N, FWD, false, 1, 8>(P, W, n1, n3, sg, nsm, st)
                     : reg_launch_y1<N, FWD, false, 0, 8>(P, W, n1, n3, sg, nsm, st);
  }
  if (RegSched<N / 2>::T >= 32 && wide)
    return kc == 2 ? reg_launch_y1<N, FWD, true, 2>(P, W, n1, n3, sg, nsm, st)
         : kc == 1 ? reg_launch_y1<N, FWD, true, 1>(P, W, n1, n3, sg, nsm, st)
                   : reg_launch_y1<N, FWD, true, 0>(P, W, n1, n3, sg, nsm, st);
  return kc == 2 ? reg_launch_y1<N, FWD, false, 2>(P, W, n1, n3, sg, nsm, st)
       : kc == 1 ? reg_launch_y1<N, FWD, false, 1>(P, W, n1, n3, sg, nsm, st)
                 : reg_launch_y1<N, FWD, false, 0>(P, W, n1, n3, sg, nsm, st);
}

}  // namespace fb
