// reg_fft.cuh -- batched 1-D real transforms with the line held in REGISTERS (power-of-two lengths 32..2048).
//
// Replaces, for the pressure solver only, what the reference gets from FFTW r2r plans
// (src/fft.f90:75-86,113-124 -> dfftw_execute_r2r, :181-193) and, on its GPU path, from cuFFT D2Z/Z2D
// plus the separate Makhoul pre/post sweeps (src/fft.f90:294-887).  Second-generation kernels: the
// shared-memory tile kernels (tile_fft.cuh / fft_p2.cuh) move every element through shared memory five
// times and spend half their instructions on tile addressing; here
//
//   * a length-N real line is one length-M = N/2 complex Stockham FFT; T = M/16 threads own a line, each
//     holds 16 complex values in registers from the global load to the global store;
//   * passes are radix-16 (4x4 in registers) followed by one radix-2/4/8 pass when M is not a power of 16;
//     between passes the T threads exchange through a small shared-memory buffer (write scattered, read
//     position j + T*u -- the same set in every pass), so M = 256 needs ONE exchange, M = 512..2048 two;
//   * the real split / Makhoul post-twiddle (forward) and merge / pre-twiddle (backward) are computed per
//     mode k from Z_k and Z_{M-k}.  M = 512, 1024: the small-radix pass runs on symmetric butterfly pairs so that
//     a thread holds both (reg_pair_pass_split / reg_pair_merge_pass; the backward line then runs the
//     transposed schedule, reg_fft_passes_T_tail) -- no further exchange.  Other lengths, and the 8-values
//     schedule, fetch Z_{M-k} through one more exchange (reg_scatter_modes + reg_split / reg_merge);
//   * the spectrum is left in natural order, interleaved: row 2k = Re X_k, row 2k+1 = Im X_k
//     (row 0 = X_0, row 1 = X_M); `reg_mode_index` tells the host which FFTW mode sits in which row so that
//     lambdaxy is permuted once at plan time (initsolver.f90:136-139 order -> this order).
//
// Host-compilable (tests/emulate): every function is per-thread, phase boundaries are where the kernels
// synchronise.
#pragma once
#include "tile_fft.cuh"

namespace fb {

enum { RF_MAXPASS = 3 };

// Device-visible description of one transform (tables live in global memory; pass twiddles are staged in
// shared memory by the kernels).
struct RegPlan {
  int N, M, kind;
  const cpx* tw[RF_MAXPASS];   // pass q > 0: tw[q][(t-1)*Ns + k] = exp(-2 pi i t k / (Ns r)),  t = 1..r-1, k = 0..Ns-1
  int tw_count[RF_MAXPASS];
  const cpx* tw8[RF_MAXPASS];  // same tables for the 8-values-per-thread schedule (lengths served by it, else null)
  const cpx* wN;               // [M]    exp(-2 pi i k / N)
  const cpx* wQ;               // [M+1]  exp(-i pi k / (2N))      (NN/DD only)
};

// compile-time schedule for a complex length M (power of two, 16 <= M <= 2048): RR complex values per thread,
// T = M / RR threads per line, radix-RR passes and one pass of radix R0 <= RR.
//   RR = 16 (default): M = 16: {16}; M <= 256: {16, M/16}; else {16, 16, M/256}   (FB_SCHED_SMALL_FIRST: small radix first)
//   RR = 8  (twice the threads at half the registers; used where it has the same number of passes: M = 512 = 8 x 8 x 8)
#ifndef FB_SCHED_SMALL_FIRST
#define FB_SCHED_SMALL_FIRST 0
#endif
template <int M_, int RR = 16>
struct RegSched {
  static constexpr int M = M_;
  static constexpr int R = RR;                         // complex values per thread
  static constexpr int T = M / R;                      // threads per line
  static constexpr int NP = (M == RR) ? 1 : (M <= RR * RR) ? 2 : 3;
  static constexpr int R0 = (NP == 1) ? RR : (NP == 2) ? M / RR : M / (RR * RR);
  static_assert(R0 >= 1 && R0 <= RR && T * R == M, "schedule not representable");
#if FB_SCHED_SMALL_FIRST
  static FB_CX int radix(int q) { return q == 0 ? R0 : RR; }
  static FB_CX int ns(int q) { return q == 0 ? 1 : (q == 1 ? R0 : R0 * RR); }
#else
  // the small radix goes LAST: its results stay in registers, so every exchange store has the radix-RR stride the
  // buffer padding is made for (a radix-2 first pass stores at a stride of two slots: two-way bank conflicts), and the
  // last pass has fewer twiddle products than a radix-RR one ((R0-1)/R0 instead of (RR-1)/RR per element)
  static FB_CX int radix(int q) { return q == NP - 1 ? R0 : RR; }
  static FB_CX int ns(int q) { return q == 0 ? 1 : (q == 1 ? RR : RR * RR); }
#endif
};

FB_CX bool reg_fft_supported(int N) {
  return N == 32 || N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048;
}

// which FFTW mode (index into the reference's eigenvalue array, 0-based) sits in spectral row r
FB_HD int reg_mode_index(int N, int kind, int r) {
  const int M = N / 2, k = r >> 1;
  if (kind_is_iv(kind)) return (r & 1) ? N - 1 - 2 * k : 2 * k;     // rows (Y_{2k}, Y_{N-1-2k})
  int q = (r & 1) ? ((k == 0) ? M : N - k) : k;
  if (kind == KIND_DD) q = N - 1 - q;
  return q;
}

// radix-16 butterfly, natural order in and out: y_t = sum_m u_m exp(SIGN 2 pi i m t / 16)
template <int SIGN> FB_HD void bfly16(double* re, double* im) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
  // step 1: for each b, radix-4 over a on elements (b, b+4, b+8, b+12): slot 4c+b <- Y_b[c]
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int b = 0; b < 4; ++b) {
    double tr[4] = { re[b], re[b + 4], re[b + 8], re[b + 12] }, ti[4] = { im[b], im[b + 4], im[b + 8], im[b + 12] };
    bfly4<SIGN>(tr, ti);
    re[b] = tr[0]; re[b + 4] = tr[1]; re[b + 8] = tr[2]; re[b + 12] = tr[3];
    im[b] = ti[0]; im[b + 4] = ti[1]; im[b + 8] = ti[2]; im[b + 12] = ti[3];
  }
  // step 2: slot 4c+b *= w16^(b c)   (SIGN = +1: conjugate)
  auto mul = [&](int i, double wr, double wi) {
    const double xr = re[i] * wr - im[i] * (SIGN * -wi), xi = re[i] * (SIGN * -wi) + im[i] * wr;
    re[i] = xr; im[i] = xi;
  };
  // w16^1 = (c1,-s1)  w16^2 = (h,-h)  w16^3 = (s1,-c1)  w16^4 = (0,-1)  w16^6 = (-h,-h)  w16^9 = (-c1,s1)
  mul(4 * 1 + 1, c1, -s1);  mul(4 * 1 + 2, h, -h);    mul(4 * 1 + 3, s1, -c1);
  mul(4 * 2 + 1, h, -h);
  {                                                    // w16^4 = -i (fwd), +i (bwd): (x + iy)(-i) = y - ix
    const int i = 4 * 2 + 2;
    const double nr = (SIGN < 0) ? im[i] : -im[i], ni = (SIGN < 0) ? -re[i] : re[i];
    re[i] = nr; im[i] = ni;
  }
  mul(4 * 2 + 3, -h, -h);
  mul(4 * 3 + 1, s1, -c1);  mul(4 * 3 + 2, -h, -h);   mul(4 * 3 + 3, -c1, s1);
  // step 3: for each c, radix-4 over b on the contiguous group 4c..4c+3: slot 4c+d <- X[c + 4d]
  double outr[16], outi[16];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int c = 0; c < 4; ++c) {
    double tr[4] = { re[4 * c], re[4 * c + 1], re[4 * c + 2], re[4 * c + 3] };
    double ti[4] = { im[4 * c], im[4 * c + 1], im[4 * c + 2], im[4 * c + 3] };
    bfly4<SIGN>(tr, ti);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int d = 0; d < 4; ++d) { outr[c + 4 * d] = tr[d]; outi[c + 4 * d] = ti[d]; }
  }
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < 16; ++i) { re[i] = outr[i]; im[i] = outi[i]; }
}

template <int R, int SIGN> FB_HD void rbfly(double* re, double* im) {
  if (R == 16) bfly16<SIGN>(re, im);
  else bfly<R, SIGN>(re, im);
}

// ---- exchange-buffer addressing ------------------------------------------------------------------
// Slot `pos` of a line's exchange buffer lives at padded index pos + (pos >> 4) (one pad slot per 16: the
// scattered 16-slot strides of the Stockham passes then hit all banks).  Every access of the passes has the form
// pos = (run-time base) + (compile-time offset c) where adding c never carries across a multiple of 16 beyond
// c's own multiples of 16 (shown case by case below), so pad(pos) = pad(base) + pad(c): the kernels compute one
// padded base per butterfly and the offsets fold into the immediate field of the LDS/STS instructions (the v7
// profile had 18 integer instructions per point, mostly this index arithmetic).  An exchange buffer XB provides
//   int  base(pos)               : scaled padded index of a run-time position
//   void st(base, coff, re, im)  : store at base + coff, coff = rf_padoff(c) (already padded, unscaled)
//   void ld(base, coff, re, im)
FB_CX int rf_pad(int pos) { return pos + (pos >> 4); }
FB_CX int rf_padoff(int c) { return c >= 0 ? c + (c >> 4) : -((-c) + ((-c) >> 4)); }
// The same contract with an XOR swizzle instead of pad slots: slot pos lives at pos ^ ((pos >> 4) & 7).  A warp's 32
// consecutive 16-byte slots then stay inside one aligned 512-byte window (four shared-memory wavefronts; with a pad slot
// in the middle they straddle five), and the 16-slot strides still spread over all eight 16-byte bank groups.  In every
// access of the passes base and offset occupy disjoint bits up to the swizzle field (the no-carry cases above), so
// swz(base + c) = swz(base) ^ swz(|c|): one LOP3 with an immediate per access.  An exchange buffer XB provides
//   static int off(c)            : the offset code of a compile-time c (rf_padoff(c) or rf_swz(|c|))
// and combines base and offset itself (+ or ^).
FB_CX int rf_swz(int pos) { return pos ^ ((pos >> 4) & 7); }
FB_CX int rf_swzoff(int c) { return rf_swz(c >= 0 ? c : -c); }

// One Stockham pass of the T threads owning a line.  Thread j holds element j + T*u in (re[u], im[u]).
// Not the last pass: results go to the exchange buffer (scattered), the caller synchronises and gathers.
// Last pass: results stay in registers, again as element (= mode) j + T*u.
template <class S, int Q, int SIGN, class XB>
FB_HD void reg_pass(double* re, double* im, int j, const cpx* tw, const XB& xb) {
  constexpr int r = S::radix(Q), Ns = S::ns(Q), NB = S::R / r, T = S::T;
  constexpr bool last = (Q == S::NP - 1);
  // pass 0 (Ns = 1): pos = jb r + t = j r + (T r) b + t with t < r | 16: no carry from t; T r is a multiple of 16 for
  // M >= 64, then the b term is a compile-time offset too.  Later passes (r = 16, one butterfly per thread):
  // pos = (j - k) 16 + k + t Ns, k < Ns: for Ns < 16 the low four bits are k + (t Ns mod 16) < 16, for Ns >= 16 they are k's.
  constexpr bool BCONST = (Ns == 1) && ((T * r) % 16 == 0);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int b = 0; b < NB; ++b) {
    const int jb = j + T * b;
    const int k = jb & (Ns - 1);
    double vr[r], vi[r];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) { vr[t] = re[b + t * NB]; vi[t] = im[b + t * NB]; }
    if (Ns > 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 1; t < r; ++t) {
        const cpx w = tw[(t - 1) * Ns + k];
        const double wy = (SIGN < 0) ? w.y : -w.y;
        const double xr = vr[t] * w.x - vi[t] * wy, xi = vr[t] * wy + vi[t] * w.x;
        vr[t] = xr; vi[t] = xi;
      }
    }
    rbfly<r, SIGN>(vr, vi);
    if (last) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 0; t < r; ++t) { re[b + t * NB] = vr[t]; im[b + t * NB] = vi[t]; }
    } else {
      const int base = BCONST ? xb.base(j * r) : xb.base((jb - k) * r + k);     // (jb / Ns) * Ns * r + k
      static_assert(NB == 1 || last || !XB::XOR, "swizzled buffers: non-final passes have one butterfly per thread");
      const int boff = BCONST ? XB::off(T * r * b) : 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 0; t < r; ++t) xb.st(base, boff + XB::off(t * Ns), vr[t], vi[t]);
    }
  }
}

// positions j + T u: for T >= 16 the offset is a multiple of 16, for T < 16 the low bits are j + (T u mod 16) < 16
template <class S, class XB>
FB_HD void reg_gather(double* re, double* im, int j, const XB& xb) {
  const int base = xb.base(j);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) xb.ld(base, XB::off(S::T * u), re[u], im[u]);
}

template <class S, class XB>
FB_HD void reg_scatter_modes(const double* re, const double* im, int j, const XB& xb) {
  const int base = xb.base(j);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) xb.st(base, XB::off(S::T * u), re[u], im[u]);
}

// partner mode of k = j + T u in the exchange buffer: M - k (k > 0), 0 (k = 0).  T >= 16: M - j - T u is a base
// (M - j; for j = 0 it is never dereferenced itself) minus a multiple of 16.
template <class S, class XB>
struct RegPartner {
  static constexpr int M = S::M;
  static constexpr bool CONSTOFF = (S::T % 16 == 0);
  int bm, b0;
  FB_HD RegPartner(int j, const XB& xb) : bm(CONSTOFF ? xb.base(M - j) : 0), b0(CONSTOFF ? (j == 0 ? xb.base(0) : xb.base(M - j)) : 0) {}
  FB_HD void ld(const XB& xb, int j, int u, double& r, double& i) const {
    if (CONSTOFF && XB::XOR) {
      // j > 0: M - j - T u = (T (R-1) + T - j) ^ (T u) (the u field of the base is all ones: subtracting is XOR);
      // j = 0: T ((R - u) mod R) is not an XOR of T u -> its own compile-time address, selected per lane
      const int a0 = XB::off(S::T * ((S::R - u) % S::R)), a1 = bm ^ XB::off(S::T * u);
      xb.ld(j == 0 ? a0 : a1, 0, r, i);
    } else if (CONSTOFF) {
      if (u == 0) xb.ld(b0, 0, r, i);
      else xb.ld(bm, XB::off(-S::T * u), r, i);
    } else {
      const int k = j + S::T * u;
      xb.ld(xb.base((M - k) & (M - 1)), 0, r, i);
    }
  }
};

// one mode of the forward split: (Z_k, Z_{M-k}) -> the two spectral rows of mode k.  w = wN[k]; MK: q = wQ[k] (Makhoul
// post-twiddle, NN/DD lines), wQM = wQ[M].x; k0: mode 0, whose rows are (X_0, X_M).  Same arithmetic as split_core (tile_fft.cuh).
template <bool MK>
FB_HD void reg_split_one(double zkr, double zki, double zjr, double zji, const cpx& w, const cpx& q, bool k0, double wQM,
                         double& xr, double& xi) {
  const double er = 0.5 * (zkr + zjr), ei = 0.5 * (zki - zji);
  const double orr = 0.5 * (zki + zji), oi = -0.5 * (zkr - zjr);
  const double wor = orr * w.x - oi * w.y, woi = orr * w.y + oi * w.x;
  xr = er + wor; xi = ei + woi;
  if (k0) xi = zkr - zki;                              // row 1 holds X_M = Re Z_0 - Im Z_0
  if (MK) {
    if (k0) { xr = 2.0 * xr; xi = 2.0 * wQM * xi; }
    else {
      const double tr = xr * q.x - xi * q.y, ti = xr * q.y + xi * q.x;
      xr = 2.0 * tr; xi = -2.0 * ti;
    }
  }
}

// forward split for this thread's modes k = j + T*u: (re,im)[u] <- the two spectral rows of mode k.
// The exchange buffer holds Z (all modes of the line).  wN / wQ may point to shared memory (indexed j + T u: immediates).
template <class S, bool MK, class XB>
FB_HD void reg_split(double* re, double* im, int j, const cpx* wN, const cpx* wQ, const XB& xb) {
  constexpr int M = S::M;
  const RegPartner<S, XB> pt(j, xb);
  const cpx* wNj = wN + j;
  const cpx* wQj = wQ + j;
  const double wQM = MK ? wQ[M].x : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) {
    double zjr, zji, xr, xi;
    pt.ld(xb, j, u, zjr, zji);
    const bool k0 = (u == 0) && (j == 0);
    reg_split_one<MK>(re[u], im[u], zjr, zji, wNj[S::T * u], MK ? wQj[S::T * u] : wNj[S::T * u], k0, wQM, xr, xi);
    re[u] = xr; im[u] = xi;
  }
}

// ---- forward: last (small-radix) pass on SYMMETRIC butterfly pairs, split in registers -------------------------------
// The last pass has radix r = R0 < RR and NB = RR / r butterflies per thread.  Butterfly jb (0 <= jb < Ns = M / r) turns
// buffer positions jb + t Ns into modes jb + t Ns, and the split pairs mode k with M - k, i.e. (jb, t) with (Ns - jb, r-1-t).
// A thread that runs butterflies jbA = j + T b and jbB = Ns - jbA (b < NB / 2) holds every pair itself: no exchange (and
// no barrier) between the last pass and the split -- two exchanges per line instead of three.  The one self-paired slot
// (j = 0, b = 0) runs butterflies 0 (modes t Ns pair as t <-> r - t; mode 0 is the (X_0, X_M) row) and Ns / 2 (t <-> r-1-t).
// Modes leave through out(k, xr, xi): the spectral layout (slot k = mode k) is the one of reg_split.
template <class S>
FB_CX bool reg_has_pair_pass() {
  return S::NP >= 2 && S::R0 < S::R && ((S::R / S::R0) % 2 == 0) && (S::T % 16 == 0) && !FB_SCHED_SMALL_FIRST;
}

template <class S, bool MK, class XB, class OUT>
FB_HD void reg_pair_pass_split(int j, const cpx* tw, const cpx* wN, const cpx* wQ, const XB& xb, const OUT& out) {
  constexpr int M = S::M, Q = S::NP - 1, r = S::radix(Q), Ns = S::ns(Q), NB = S::R / r, T = S::T;
  static_assert(Ns * r == M && NB % 2 == 0, "pair pass needs an even number of small-radix butterflies per thread");
  const int bA = xb.base(j), bB = xb.base(Ns - j);
  const cpx* twA = tw + j;
  const double wQM = MK ? wQ[M].x : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int b = 0; b < NB / 2; ++b) {
    const bool self = (b == 0) && (j == 0);
    // butterfly B of lane 0 sits at compile-time positions (Ns - T b is not an XOR / no-carry offset of base(Ns))
    const int jbB = (j == 0) ? (b == 0 ? Ns / 2 : Ns - T * b) : Ns - j - T * b;
    double ar[r], ai[r], br[r], bi[r];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) {
      xb.ld(bA, XB::off(T * b + t * Ns), ar[t], ai[t]);
      const int a0 = xb.base((b == 0 ? Ns / 2 : Ns - T * b) + t * Ns), a1 = xb.addr(bB, XB::offsub(t * Ns, T * b));
      xb.ld(j == 0 ? a0 : a1, 0, br[t], bi[t]);
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 1; t < r; ++t) {                      // forward twiddles w^(t k), k = jb
      const cpx wa = twA[(t - 1) * Ns + T * b], wb = tw[(t - 1) * Ns + jbB];
      const double xr = ar[t] * wa.x - ai[t] * wa.y, xi = ar[t] * wa.y + ai[t] * wa.x;
      ar[t] = xr; ai[t] = xi;
      const double yr = br[t] * wb.x - bi[t] * wb.y, yi = br[t] * wb.y + bi[t] * wb.x;
      br[t] = yr; bi[t] = yi;
    }
    rbfly<r, -1>(ar, ai);
    rbfly<r, -1>(br, bi);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) {
      // mode kA = jbA + t Ns pairs with B's output r-1-t (self: A's output (r - t) mod r);
      // mode kB = jbB + t Ns pairs with A's output r-1-t (self: B's output r-1-t)
      const double pAr = self ? ar[(r - t) % r] : br[r - 1 - t], pAi = self ? ai[(r - t) % r] : bi[r - 1 - t];
      const double pBr = self ? br[r - 1 - t] : ar[r - 1 - t], pBi = self ? bi[r - 1 - t] : ai[r - 1 - t];
      const int kA = j + T * b + t * Ns, kB = jbB + t * Ns;
      double xr, xi;
      reg_split_one<MK>(ar[t], ai[t], pAr, pAi, wN[kA], MK ? wQ[kA] : wN[kA], self && t == 0, wQM, xr, xi);
      out(kA, xr, xi);
      reg_split_one<MK>(br[t], bi[t], pBr, pBi, wN[kB], MK ? wQ[kB] : wN[kB], false, wQM, xr, xi);
      out(kB, xr, xi);
    }
  }
}

// one mode of the backward merge: the spectral rows of modes k and M - k -> Z_k.  w = wN[k]; MK: qk = wQ[k], qj = wQ[M-k];
// k0: mode 0 from its (X_0, X_M) rows.
template <bool MK>
FB_HD void reg_merge_one(double xkr, double xki, double xjr, double xji, const cpx& w, const cpx& qk, const cpx& qj, bool k0,
                         double wQM, double& zr, double& zi) {
  if (k0) {
    double xm = xki;
    if (MK) xm = 2.0 * wQM * xm;
    zr = xkr + xm; zi = xkr - xm;
    return;
  }
  if (MK) {
    const double vkr = xkr * qk.x - xki * qk.y, vki = -xki * qk.x - xkr * qk.y;
    const double vjr = xjr * qj.x - xji * qj.y, vji = -xji * qj.x - xjr * qj.y;
    xkr = vkr; xki = vki; xjr = vjr; xji = vji;
  }
  const double sr = xkr + xjr, si = xki - xji, dr = xkr - xjr, di = xki + xji;
  const double cr = dr * w.x + di * w.y, ci = di * w.x - dr * w.y;     // conj(w) D
  zr = sr - ci; zi = si + cr;                                           // S + i conj(w) D
}

template <class S, bool MK, class XB>
FB_HD void reg_merge(double* re, double* im, int j, const cpx* wN, const cpx* wQ, const XB& xb) {
  constexpr int M = S::M;
  const RegPartner<S, XB> pt(j, xb);
  const cpx* wNj = wN + j;
  const cpx* wQj = wQ + j;
  const cpx* wQm = wQ + (M - j);
  const double wQM = MK ? wQ[M].x : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) {
    double xjr, xji, zr, zi;
    pt.ld(xb, j, u, xjr, xji);
    const bool k0 = (u == 0) && (j == 0);
    const cpx w = wNj[S::T * u];
    reg_merge_one<MK>(re[u], im[u], xjr, xji, w, MK ? wQj[S::T * u] : w, MK ? wQm[-S::T * u] : w, k0, wQM, zr, zi);
    re[u] = zr; im[u] = zi;
  }
}

// ---- backward: the transposed schedule ------------------------------------------------------------------------------------
// The DFT matrix is symmetric, so the forward factorisation P_last ... P_1 P_0 read backwards with every pass transposed
// (gather <-> scatter swapped, twiddles after the butterfly) is the same transform; with conjugated twiddles, the inverse.
// Run that way the backward line STARTS with the small-radix pass on symmetric butterfly pairs, whose inputs are exactly the
// (k, M - k) pairs of the merge: rows come in through in(k, xr, xi), are merged in registers, and every exchange of the
// remaining radix-RR passes has the forward one's (conflict-free) access patterns -- two exchanges per line instead of three.
template <class S, bool MK, class XB, class IN>
FB_HD void reg_pair_merge_pass(int j, const cpx* tw, const cpx* wN, const cpx* wQ, const XB& xb, const IN& in) {
  constexpr int M = S::M, Q = S::NP - 1, r = S::radix(Q), Ns = S::ns(Q), NB = S::R / r, T = S::T;
  static_assert(Ns * r == M && NB % 2 == 0, "pair pass needs an even number of small-radix butterflies per thread");
  const int bA = xb.base(j), bB = xb.base(Ns - j);
  const cpx* twA = tw + j;
  const double wQM = MK ? wQ[M].x : 0.0;
  // every row of the line first (one round of independent loads, as the exchange-based path has), then slot by slot
  double xr_[S::R], xi_[S::R];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int b = 0; b < NB / 2; ++b) {
    const int jbB = (j == 0) ? (b == 0 ? Ns / 2 : Ns - T * b) : Ns - j - T * b;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) {
      in(j + T * b + t * Ns, xr_[2 * r * b + t], xi_[2 * r * b + t]);
      in(jbB + t * Ns, xr_[2 * r * b + r + t], xi_[2 * r * b + r + t]);
    }
  }
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int b = 0; b < NB / 2; ++b) {
    const bool self = (b == 0) && (j == 0);
    const int jbB = (j == 0) ? (b == 0 ? Ns / 2 : Ns - T * b) : Ns - j - T * b;
    const double* xar = xr_ + 2 * r * b; const double* xai = xi_ + 2 * r * b;
    const double* xbr = xar + r; const double* xbi = xai + r;
    double ar[r], ai[r], br[r], bi[r];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) {                      // partner rows: see reg_pair_pass_split
      const double pAr = self ? xar[(r - t) % r] : xbr[r - 1 - t], pAi = self ? xai[(r - t) % r] : xbi[r - 1 - t];
      const double pBr = self ? xbr[r - 1 - t] : xar[r - 1 - t], pBi = self ? xbi[r - 1 - t] : xai[r - 1 - t];
      const int kA = j + T * b + t * Ns, kB = jbB + t * Ns;
      const cpx wa = wN[kA], wb = wN[kB];
      reg_merge_one<MK>(xar[t], xai[t], pAr, pAi, wa, MK ? wQ[kA] : wa, MK ? wQ[M - kA] : wa, self && t == 0, wQM, ar[t], ai[t]);
      reg_merge_one<MK>(xbr[t], xbi[t], pBr, pBi, wb, MK ? wQ[kB] : wb, MK ? wQ[M - kB] : wb, false, wQM, br[t], bi[t]);
    }
    rbfly<r, +1>(ar, ai);
    rbfly<r, +1>(br, bi);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 1; t < r; ++t) {                      // conjugate twiddles w^(t k), k = jb, AFTER the butterfly
      const cpx wa = twA[(t - 1) * Ns + T * b], wb = tw[(t - 1) * Ns + jbB];
      const double xr = ar[t] * wa.x + ai[t] * wa.y, xi = ai[t] * wa.x - ar[t] * wa.y;
      ar[t] = xr; ai[t] = xi;
      const double yr = br[t] * wb.x + bi[t] * wb.y, yi = bi[t] * wb.x - br[t] * wb.y;
      br[t] = yr; bi[t] = yi;
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) {
      xb.st(bA, XB::off(T * b + t * Ns), ar[t], ai[t]);
      const int a0 = xb.base((b == 0 ? Ns / 2 : Ns - T * b) + t * Ns), a1 = xb.addr(bB, XB::offsub(t * Ns, T * b));
      xb.st(j == 0 ? a0 : a1, 0, br[t], bi[t]);
    }
  }
}

// transposed radix-RR pass Q (one butterfly per thread), first half: read where the forward pass stores
template <class S, int Q, class XB>
FB_HD void reg_pass_T_load(double* re, double* im, int j, const XB& xb) {
  constexpr int r = S::radix(Q), Ns = S::ns(Q), T = S::T;
  static_assert(S::R / r == 1, "transposed passes: radix RR only");
  constexpr bool BCONST = (Ns == 1) && ((T * r) % 16 == 0);
  const int k = j & (Ns - 1);
  const int base = BCONST ? xb.base(j * r) : xb.base((j - k) * r + k);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int t = 0; t < r; ++t) xb.ld(base, XB::off(t * Ns), re[t], im[t]);
}
// second half (after the caller's barrier): butterfly, twiddles, and -- unless it is pass 0, whose results are the line's
// elements j + T t and stay in registers -- store where the forward pass gathers
template <class S, int Q, int SIGN, class XB>
FB_HD void reg_pass_T_finish(double* re, double* im, int j, const cpx* tw, const XB& xb) {
  constexpr int r = S::radix(Q), Ns = S::ns(Q), T = S::T;
  const int k = j & (Ns - 1);
  rbfly<r, SIGN>(re, im);
  if (Ns > 1) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 1; t < r; ++t) {
      const cpx w = tw[(t - 1) * Ns + k];
      const double wy = (SIGN < 0) ? w.y : -w.y;
      const double xr = re[t] * w.x - im[t] * wy, xi = re[t] * wy + im[t] * w.x;
      re[t] = xr; im[t] = xi;
    }
  }
  if (Q > 0) {
    const int bs = xb.base(j);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < r; ++t) xb.st(bs, XB::off(T * t), re[t], im[t]);
  }
}
// the transposed passes after reg_pair_merge_pass; on return (re, im)[u] = element j + T u of the inverse transform
template <class S, int SIGN, class XB, class SYNC>
FB_HD void reg_fft_passes_T_tail(double* re, double* im, int j, const cpx* const* tw, const XB& xb, const SYNC& sync) {
  static_assert(S::NP >= 2, "no tail pass");
  sync();
  if constexpr (S::NP > 2) {
    reg_pass_T_load<S, 1>(re, im, j, xb);
    sync();
    reg_pass_T_finish<S, 1, SIGN>(re, im, j, tw[1], xb);
    sync();
  }
  reg_pass_T_load<S, 0>(re, im, j, xb);
  reg_pass_T_finish<S, 0, SIGN>(re, im, j, tw[0], xb);
}

// ---- types IV (ND / DN: REDFT11 / RODFT11, the same transform in both directions; tile_fft.cuh has the algebra) ----
// pre-twiddle of this thread's packed elements m = j + T u (wQ[m] = e^{-i pi (4m+1)/(4N)}); CONJ: the inverse-sign passes
// that follow then compute the conjugate of the forward transform
template <class S, bool CONJ>
FB_HD void reg_iv_pre(double* re, double* im, int j, const cpx* wQ) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) {
    const cpx w = wQ[j + S::T * u];
    const double vr = re[u] * w.x - im[u] * w.y, vi = re[u] * w.y + im[u] * w.x;
    re[u] = vr; im[u] = CONJ ? -vi : vi;
  }
}
// post-twiddle of this thread's modes k = j + T u (wN[k] = e^{-i pi k/N}): (a, b) = (2 Re w, -2 Im w).
// Forward: rows (2k, 2k+1) <- (Y_{2k}, Y_{N-1-2k}) = (a, b), DN: (a, -b).  Backward (CONJ input): the physical pair of
// packed element k, ND: (x_{2k}, x_{N-1-2k}) = (a, b); DN (reversed line): (x_{N-1-2k}, x_{2k}) = (-b, a).
template <class S, bool FWD>
FB_HD void reg_iv_post(double* re, double* im, int j, const cpx* wN, bool dn) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int u = 0; u < S::R; ++u) {
    const double vr = re[u], vi = FWD ? im[u] : -im[u];
    const cpx w = wN[j + S::T * u];
    const double a = 2.0 * (vr * w.x - vi * w.y), b = -2.0 * (vr * w.y + vi * w.x);
    if (!dn) { re[u] = a; im[u] = b; }
    else if (FWD) { re[u] = a; im[u] = -b; }
    else { re[u] = -b; im[u] = a; }
  }
}

// Physical rows of packed element m = j + T u of a Makhoul (NN/DD) line, T = N/(2 RR) threads per line, RR values each:
//   u <  RR/2 (m <  N/4): e0 = 4 m,            e1 = e0 + 2      (even elements; DD sign +)
//   u >= RR/2 (m >= N/4): e0 = 2 N - 1 - 4 m,  e1 = e0 - 2      (odd elements;  DD sign -)
// (slot_to_elem, tile_fft.cuh, for v = 2m and 2m+1.)  So a thread needs two row bases, 4 j and 2N-1-4j, and
// compile-time multiples of the row stride.
template <int N, int RR = 16>
struct MkRows {
  static constexpr int T = N / (2 * RR);
  static FB_CX bool upper(int u) { return u >= RR / 2; }
  static FB_CX int off0(int u) { return u < RR / 2 ? 4 * T * u : -4 * T * u; }
  static FB_CX int off1(int u) { return u < RR / 2 ? 4 * T * u + 2 : -4 * T * u - 2; }
  FB_HD static int base_lo(int j) { return 4 * j; }
  FB_HD static int base_hi(int j) { return 2 * N - 1 - 4 * j; }
};

// physical element indices (and DST sign) of packed complex element m: z_m = s0 x[e0] + i s1 x[e1]
FB_HD void reg_phys_slots(int kind, int N, int m, int& e0, int& e1, double& s0, double& s1) {
  slot_to_elem(kind, N, 2 * m, e0, s0);
  slot_to_elem(kind, N, 2 * m + 1, e1, s1);
}

// complete forward / backward passes between gather points; SYNC is the caller's barrier functor
// all passes but the last one; the last of them leaves its results in the exchange buffer (reg_pair_pass_split follows
// after the caller's barrier)
template <class S, int SIGN, class XB, class SYNC>
FB_HD void reg_fft_passes_head(double* re, double* im, int j, const cpx* const* tw, const XB& xb, const SYNC& sync) {
  static_assert(S::NP >= 2, "no head pass");
  reg_pass<S, 0, SIGN>(re, im, j, tw[0], xb);
  if constexpr (S::NP > 2) {
    sync(); reg_gather<S>(re, im, j, xb); sync();
    reg_pass<S, 1, SIGN>(re, im, j, tw[1], xb);
  }
}

template <class S, int SIGN, class XB, class SYNC>
FB_HD void reg_fft_passes(double* re, double* im, int j, const cpx* const* tw, const XB& xb, const SYNC& sync) {
  reg_pass<S, 0, SIGN>(re, im, j, tw[0], xb);
  if constexpr (S::NP > 1) {
    sync(); reg_gather<S>(re, im, j, xb); sync();
    reg_pass<S, 1, SIGN>(re, im, j, tw[1], xb);
  }
  if constexpr (S::NP > 2) {
    sync(); reg_gather<S>(re, im, j, xb); sync();
    reg_pass<S, 2, SIGN>(re, im, j, tw[2], xb);
  }
}

}  // namespace fb
