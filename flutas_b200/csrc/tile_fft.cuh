// tile_fft.cuh -- batched 1-D real transforms on a shared-memory tile, written once for x and y lines.
//
// Replaces, for the pressure solver only, what the reference gets from FFTW r2r plans
// (src/fft.f90:75-86,113-124 -> dfftw_execute_r2r, :181-193) and, on its GPU path, from cuFFT D2Z/Z2D
// plus the separate Makhoul pre/post "signal processing" sweeps (src/fft.f90:294-887).
//
// A tile holds TB lines of N reals as [row][lane]: lane = line index (fastest), row = element.  Every
// butterfly/post-processing access is "TB consecutive doubles of one row", so lanes never bank-conflict
// and all lanes share one twiddle (broadcast).  A length-N real transform is one length-M=N/2 complex
// FFT (split re/im planes: row m = Re z_m, row M+m = Im z_m) plus a fused split/Makhoul step:
//
//   forward : load (PP: v=e | NN: Makhoul even/odd permutation | DD: same + sign of odd e)
//             -> in-place DIF passes (natural -> digit-reversed) -> split (+ e^{-i pi k/2N} twiddle)
//   backward: inverse twiddle/split -> in-place DIT passes (digit-reversed -> natural) -> un-permute store
//
// The spectral layout is whatever the forward leaves in the tile (digit-reversed, re/im planes); it
// is never reordered: `mode_index()` in plan.h tells the host which FFTW mode sits in which row so
// that lambdaxy is permuted once at plan time ("physical order free").
//
// The same code compiles for the host (tests/emulate) so index logic is testable without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#define FB_CX __host__ __device__ constexpr
#else
#define FB_HD inline
#define FB_CX constexpr
#endif

namespace fb {

struct cpx { double x, y; };

// PP: R2HC/HC2R; NN: REDFT10/01, DD: RODFT10/01 (Makhoul through the real FFT); ND: REDFT11, DN: RODFT11 (types IV:
// one transform for both directions, src/fft.f90:256-263), computed as  z_m = x_{2m} + i x_{N-1-2m},
// v_m = z_m e^{-i pi (4m+1)/(4N)}, V = FFT_{N/2}(v), w_k = V_k e^{-i pi k/N}, Y_{2k} = 2 Re w_k, Y_{N-1-2k} = -2 Im w_k;
// RODFT11(x)_k = (-1)^k REDFT11(reversed x)_k.
enum LineKind { KIND_PP = 0, KIND_NN = 1, KIND_DD = 2, KIND_ND = 3, KIND_DN = 4 };
FB_CX bool kind_is_iv(int kind) { return kind == KIND_ND || kind == KIND_DN; }
FB_CX bool kind_is_makhoul(int kind) { return kind == KIND_NN || kind == KIND_DD; }
enum { FB_MAX_PASS = 12 };

// Device-visible description of one transform length/kind (tables live in global memory).
struct LinePlan {
  int N;                    // real length (even)
  int M;                    // complex length N/2
  int kind;                 // LineKind
  int npass;
  int radix[FB_MAX_PASS];   // DIF pass q uses radix[q]
  int sub[FB_MAX_PASS];     // butterfly stride of pass q: M / (radix[0]*...*radix[q])
  const cpx* wM;            // [M]      exp(-2 pi i k / M)
  const cpx* wN;            // [M/2+1]  exp(-2 pi i k / N)
  const cpx* wQ;            // [M+1]    exp(-i pi k / (2N))      (NN/DD only)
  const int* pos;           // [M]      row of complex mode k after the forward passes
};

// ---------------------------------------------------------------------------------------------
// tile addressing.  ROT=false: column == lane (y lines, loaded lane-contiguous from global).
// ROT=true : column rotated by a row-dependent amount so that the transposing x-line load/store
// (a warp writes 32 consecutive elements of ONE line, i.e. 32 different rows) is conflict-free.
template <int TB>
struct TileShape {
  static constexpr int SH = (TB >= 16) ? 0 : (TB == 8) ? 1 : (TB == 4) ? 2 : 3;   // log2(16/TB)
};

template <int TB, bool ROT>
FB_HD int taddr(int m, int part, int M, int lane) {
  const int row = m + part * M;
  if (ROT) {
    const int rot = ((m >> TileShape<TB>::SH) & (TB / 2 - 1)) | (part * (TB / 2));
    return row * TB + ((lane + rot) & (TB - 1));
  }
  return row * TB + lane;
}

// element e of a physical line -> (row m, plane part, sign) of the packed complex sequence
FB_HD void elem_to_slot(int kind, int N, int e, int& m, int& part, double& sgn) {
  int v = e;
  sgn = 1.0;
  if (kind_is_iv(kind)) {                      // type IV: z_m = x_{2m} + i x_{N-1-2m}; DN: the line reversed
    const bool odd = (e & 1);
    m = odd ? (N - 1 - e) >> 1 : e >> 1;
    part = (odd != (kind == KIND_DN)) ? 1 : 0;
    return;
  }
  if (kind != KIND_PP) {                       // Makhoul: v(n)=x(2n), v(N-1-n)=x(2n+1)  (fft.f90:431-444)
    v = (e & 1) ? (N - 1 - (e >> 1)) : (e >> 1);
    if (kind == KIND_DD && (e & 1)) sgn = -1.0;   // DST-II/III via sign flip of odd inputs (fft.f90:417-428,859-875)
  }
  m = v >> 1;
  part = v & 1;
}

// ---------------------------------------------------------------------------------------------
// radix butterflies: y_t = sum_m u_m exp(SIGN * 2 pi i m t / R), in place on (re[], im[])
template <int SIGN> FB_HD void bfly2(double* re, double* im) {
  const double ar = re[0], ai = im[0];
  re[0] = ar + re[1]; im[0] = ai + im[1];
  re[1] = ar - re[1]; im[1] = ai - im[1];
}

template <int SIGN> FB_HD void bfly3(double* re, double* im) {
  const double c = -0.5, s = SIGN * 0.86602540378443864676;   // exp(SIGN 2 pi i/3) = c + i s
  const double sr = re[1] + re[2], si = im[1] + im[2];
  const double dr = re[1] - re[2], di = im[1] - im[2];
  const double tr = re[0] + c * sr, ti = im[0] + c * si;
  re[0] += sr; im[0] += si;
  re[1] = tr - s * di; im[1] = ti + s * dr;
  re[2] = tr + s * di; im[2] = ti - s * dr;
}

template <int SIGN> FB_HD void bfly4(double* re, double* im) {
  const double s02r = re[0] + re[2], s02i = im[0] + im[2], d02r = re[0] - re[2], d02i = im[0] - im[2];
  const double s13r = re[1] + re[3], s13i = im[1] + im[3], d13r = re[1] - re[3], d13i = im[1] - im[3];
  // (SIGN i) * d13
  const double jr = -SIGN * d13i, ji = SIGN * d13r;
  re[0] = s02r + s13r; im[0] = s02i + s13i;
  re[1] = d02r + jr;   im[1] = d02i + ji;
  re[2] = s02r - s13r; im[2] = s02i - s13i;
  re[3] = d02r - jr;   im[3] = d02i - ji;
}

template <int SIGN> FB_HD void bfly5(double* re, double* im) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
  const double s1 = SIGN * 0.95105651629515357212, s2 = SIGN * 0.58778525229247312917;
  const double a1r = re[1] + re[4], a1i = im[1] + im[4], b1r = re[1] - re[4], b1i = im[1] - im[4];
  const double a2r = re[2] + re[3], a2i = im[2] + im[3], b2r = re[2] - re[3], b2i = im[2] - im[3];
  const double x0r = re[0], x0i = im[0];
  const double p1r = x0r + c1 * a1r + c2 * a2r, p1i = x0i + c1 * a1i + c2 * a2i;
  const double p2r = x0r + c2 * a1r + c1 * a2r, p2i = x0i + c2 * a1i + c1 * a2i;
  const double q1r = s1 * b1r + s2 * b2r, q1i = s1 * b1i + s2 * b2i;   // multiplied by i below
  const double q2r = s2 * b1r - s1 * b2r, q2i = s2 * b1i - s1 * b2i;
  re[0] = x0r + a1r + a2r; im[0] = x0i + a1i + a2i;
  re[1] = p1r - q1i; im[1] = p1i + q1r;
  re[4] = p1r + q1i; im[4] = p1i - q1r;
  re[2] = p2r - q2i; im[2] = p2i + q2r;
  re[3] = p2r + q2i; im[3] = p2i - q2r;
}

template <int SIGN> FB_HD void bfly8(double* re, double* im) {
  const double h = 0.70710678118654752440;
  double er[4] = { re[0], re[2], re[4], re[6] }, ei[4] = { im[0], im[2], im[4], im[6] };
  double qr[4] = { re[1], re[3], re[5], re[7] }, qi[4] = { im[1], im[3], im[5], im[7] };
  bfly4<SIGN>(er, ei);
  bfly4<SIGN>(qr, qi);
  // twiddles exp(SIGN 2 pi i t/8), t = 1,2,3
  double tr, ti;
  tr = h * (qr[1] - SIGN * qi[1]); ti = h * (qi[1] + SIGN * qr[1]); qr[1] = tr; qi[1] = ti;
  tr = -SIGN * qi[2]; ti = SIGN * qr[2]; qr[2] = tr; qi[2] = ti;
  tr = h * (-qr[3] - SIGN * qi[3]); ti = h * (-qi[3] + SIGN * qr[3]); qr[3] = tr; qi[3] = ti;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int t = 0; t < 4; ++t) {
    re[t] = er[t] + qr[t]; im[t] = ei[t] + qi[t];
    re[t + 4] = er[t] - qr[t]; im[t + 4] = ei[t] - qi[t];
  }
}

template <int R, int SIGN> FB_HD void bfly(double* re, double* im) {
  if (R == 2) bfly2<SIGN>(re, im);
  else if (R == 3) bfly3<SIGN>(re, im);
  else if (R == 4) bfly4<SIGN>(re, im);
  else if (R == 5) bfly5<SIGN>(re, im);
  else bfly8<SIGN>(re, im);
}

// ---------------------------------------------------------------------------------------------
// accessors: where a pass / split step reads and writes complex element m (re, im).
//   TileAcc  : the shared-memory tile.
//   LineAcc  : a strided line in global memory (y lines: element e at base[e*stride]), with the
//              physical <-> packed mapping (Makhoul permutation, DST sign) applied on the fly; used to
//              fuse the first pass with the global load and the last pass with the global store.
//   SpecAcc  : the spectral layout in global memory (row r at base[r*stride]), for the fused split/merge.
template <int TB, bool ROT>
struct TileAcc {
  double* tile; int M, lane;
  FB_HD void ld(int m, double& re, double& im) const {
    re = tile[taddr<TB, ROT>(m, 0, M, lane)];
    im = tile[taddr<TB, ROT>(m, 1, M, lane)];
  }
  FB_HD void st(int m, double re, double im) const {
    tile[taddr<TB, ROT>(m, 0, M, lane)] = re;
    tile[taddr<TB, ROT>(m, 1, M, lane)] = im;
  }
};

// packed index v -> physical element e and sign (inverse of elem_to_slot)
FB_HD void slot_to_elem(int kind, int N, int v, int& e, double& sgn) {
  e = v;
  sgn = 1.0;
  if (kind_is_iv(kind)) {
    const int m = v >> 1;
    const bool second = ((v & 1) != 0) != (kind == KIND_DN);
    e = second ? N - 1 - 2 * m : 2 * m;
    return;
  }
  if (kind != KIND_PP) {
    e = (2 * v < N) ? 2 * v : 2 * (N - 1 - v) + 1;
    if (kind == KIND_DD && (e & 1)) sgn = -1.0;
  }
}

struct LineAcc {
  double* base; long stride; int kind, N; bool live; double scale;
  FB_HD void ld(int m, double& re, double& im) const {
    int e0, e1; double s0, s1;
    slot_to_elem(kind, N, 2 * m, e0, s0);
    slot_to_elem(kind, N, 2 * m + 1, e1, s1);
#if defined(__CUDA_ARCH__)
    re = live ? s0 * __ldcs(base + (long)e0 * stride) : 0.0;
    im = live ? s1 * __ldcs(base + (long)e1 * stride) : 0.0;
#else
    re = live ? s0 * base[(long)e0 * stride] : 0.0;
    im = live ? s1 * base[(long)e1 * stride] : 0.0;
#endif
  }
  FB_HD void st(int m, double re, double im) const {
    if (!live) return;
    int e0, e1; double s0, s1;
    slot_to_elem(kind, N, 2 * m, e0, s0);
    slot_to_elem(kind, N, 2 * m + 1, e1, s1);
#if defined(__CUDA_ARCH__)
    __stcs(base + (long)e0 * stride, s0 * scale * re);
    __stcs(base + (long)e1 * stride, s1 * scale * im);
#else
    base[(long)e0 * stride] = s0 * scale * re;
    base[(long)e1 * stride] = s1 * scale * im;
#endif
  }
};

struct SpecAcc {
  double* base; long stride; int M; bool live;
  FB_HD void ld(int m, double& re, double& im) const {
#if defined(__CUDA_ARCH__)
    re = live ? __ldcs(base + (long)m * stride) : 0.0;
    im = live ? __ldcs(base + (long)(M + m) * stride) : 0.0;
#else
    re = live ? base[(long)m * stride] : 0.0;
    im = live ? base[(long)(M + m) * stride] : 0.0;
#endif
  }
  FB_HD void st(int m, double re, double im) const {
    if (!live) return;
#if defined(__CUDA_ARCH__)
    __stcs(base + (long)m * stride, re);
    __stcs(base + (long)(M + m) * stride, im);
#else
    base[(long)m * stride] = re;
    base[(long)(M + m) * stride] = im;
#endif
  }
};

// ---------------------------------------------------------------------------------------------
// one radix-R pass.  FWD: DIF (butterfly, then twiddle).  !FWD: DIT (conj twiddle, then inverse
// butterfly) -- exactly undoes the FWD pass up to the factor R.  M and s may be compile-time constants
// at the call site (power-of-two kernels): everything here is force-inlined so they fold.
template <int R, bool FWD, class LD, class ST>
FB_HD void pass_core(int M, int s, const cpx* wM, int worker, int nworkers, const LD& src, const ST& dst) {
  const int Lc = s * R;
  const int tstep = M / Lc;
  const int nb = M / R;
  for (int b = worker; b < nb; b += nworkers) {
    const int blk = b / s, n = b - blk * s;
    const int base = blk * Lc + n;
    double re[R], im[R];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < R; ++t) src.ld(base + t * s, re[t], im[t]);
    if (FWD) {
      bfly<R, -1>(re, im);
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 1; t < R; ++t) {
        const cpx w = wM[n * t * tstep];
        const double xr = re[t] * w.x - im[t] * w.y, xi = re[t] * w.y + im[t] * w.x;
        re[t] = xr; im[t] = xi;
      }
    } else {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int t = 1; t < R; ++t) {
        const cpx w = wM[n * t * tstep];            // multiply by conj(w)
        const double xr = re[t] * w.x + im[t] * w.y, xi = im[t] * w.x - re[t] * w.y;
        re[t] = xr; im[t] = xi;
      }
      bfly<R, +1>(re, im);
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int t = 0; t < R; ++t) dst.st(base + t * s, re[t], im[t]);
  }
}

template <int TB, bool ROT, int R, bool FWD>
FB_HD void pass_R(double* tile, const LinePlan& P, const cpx* wM, int q, int lane, int worker, int nworkers) {
  const TileAcc<TB, ROT> acc{tile, P.M, lane};
  pass_core<R, FWD>(P.M, P.sub[q], wM, worker, nworkers, acc, acc);
}

template <int TB, bool ROT, bool FWD>
FB_HD void fft_pass(double* tile, const LinePlan& P, const cpx* wM, int q, int lane, int worker, int nworkers) {
  switch (P.radix[q]) {
    case 2: pass_R<TB, ROT, 2, FWD>(tile, P, wM, q, lane, worker, nworkers); break;
    case 3: pass_R<TB, ROT, 3, FWD>(tile, P, wM, q, lane, worker, nworkers); break;
    case 4: pass_R<TB, ROT, 4, FWD>(tile, P, wM, q, lane, worker, nworkers); break;
    case 5: pass_R<TB, ROT, 5, FWD>(tile, P, wM, q, lane, worker, nworkers); break;
    default: pass_R<TB, ROT, 8, FWD>(tile, P, wM, q, lane, worker, nworkers); break;
  }
}

// ---------------------------------------------------------------------------------------------
// forward split: complex FFT of the packed sequence -> spectrum of the real line.
//   PP : slot(pos k) <- (Re X_k, Im X_k), k=1..M-1 ; slot(0) <- (X_0, X_M)        [R2HC content]
//   NN : slot(pos k) <- (Y_k, Y_{N-k}) ; slot(0) <- (Y_0, Y_M)                    [REDFT10 content]
//   DD : as NN on the sign-flipped input; row holding DCT mode q holds DST mode N-1-q (fft.f90:537-560)
template <class LD, class ST>
FB_HD void split_core(int M, int kind, const cpx* wN, const cpx* wQ, const int* pos, int worker, int nworkers,
                      const LD& src, const ST& dst) {
  const bool mk = (kind != KIND_PP);
  for (int k = worker; 2 * k <= M; k += nworkers) {
    if (k == 0) {
      double zr, zi;
      src.ld(0, zr, zi);
      double x0 = zr + zi, xm = zr - zi;
      if (mk) { x0 = 2.0 * x0; xm = 2.0 * wQ[M].x * xm; }
      dst.st(0, x0, xm);
      continue;
    }
    const int pk = pos[k];
    if (2 * k == M) {                                   // X = conj(Z)
      double xr, xi;
      src.ld(pk, xr, xi);
      xi = -xi;
      if (mk) {
        const cpx q = wQ[k];
        const double tr = xr * q.x - xi * q.y, ti = xr * q.y + xi * q.x;
        xr = 2.0 * tr; xi = -2.0 * ti;
      }
      dst.st(pk, xr, xi);
      continue;
    }
    const int pj = pos[M - k];
    double zkr, zki, zjr, zji;
    src.ld(pk, zkr, zki);
    src.ld(pj, zjr, zji);
    // E = (Z_k + conj Z_j)/2 ; O = -(i/2)(Z_k - conj Z_j) ; X_k = E + w^k O ; X_j = conj(E - w^k O)
    const double er = 0.5 * (zkr + zjr), ei = 0.5 * (zki - zji);
    const double orr = 0.5 * (zki + zji), oi = -0.5 * (zkr - zjr);
    const cpx w = wN[k];
    const double wor = orr * w.x - oi * w.y, woi = orr * w.y + oi * w.x;
    double xkr = er + wor, xki = ei + woi, xjr = er - wor, xji = -(ei - woi);
    if (mk) {
      const cpx qk = wQ[k], qj = wQ[M - k];
      const double tkr = xkr * qk.x - xki * qk.y, tki = xkr * qk.y + xki * qk.x;
      const double tjr = xjr * qj.x - xji * qj.y, tji = xjr * qj.y + xji * qj.x;
      xkr = 2.0 * tkr; xki = -2.0 * tki; xjr = 2.0 * tjr; xji = -2.0 * tji;
    }
    dst.st(pk, xkr, xki);
    dst.st(pj, xjr, xji);
  }
}

// backward merge: inverse of split_core up to the FFTW scale (HC2R: N, REDFT01/RODFT01: 2N)
template <class LD, class ST>
FB_HD void merge_core(int M, int kind, const cpx* wN, const cpx* wQ, const int* pos, int worker, int nworkers,
                      const LD& src, const ST& dst) {
  const bool mk = (kind != KIND_PP);
  for (int k = worker; 2 * k <= M; k += nworkers) {
    if (k == 0) {
      double x0, xm;
      src.ld(0, x0, xm);
      if (mk) xm = 2.0 * wQ[M].x * xm;                  // V'_M = sqrt(2) Y_M
      dst.st(0, x0 + xm, x0 - xm);
      continue;
    }
    const int pk = pos[k];
    if (2 * k == M) {                                   // Z' = 2 conj(X)
      double xr, xi;
      src.ld(pk, xr, xi);
      if (mk) {                                         // V' = conj(q) (Y_k - i Y_{N-k})
        const cpx q = wQ[k];
        const double vr = xr * q.x - xi * q.y, vi = -xi * q.x - xr * q.y;
        xr = vr; xi = vi;
      }
      dst.st(pk, 2.0 * xr, -2.0 * xi);
      continue;
    }
    const int pj = pos[M - k];
    double xkr, xki, xjr, xji;
    src.ld(pk, xkr, xki);
    src.ld(pj, xjr, xji);
    if (mk) {
      const cpx qk = wQ[k], qj = wQ[M - k];
      const double vkr = xkr * qk.x - xki * qk.y, vki = -xki * qk.x - xkr * qk.y;
      const double vjr = xjr * qj.x - xji * qj.y, vji = -xji * qj.x - xjr * qj.y;
      xkr = vkr; xki = vki; xjr = vjr; xji = vji;
    }
    // S = X_k + conj X_j ; D = X_k - conj X_j ; T = i conj(w^k) D ; Z'_k = S + T ; Z'_j = conj(S - T)
    const double sr = xkr + xjr, si = xki - xji, dr = xkr - xjr, di = xki + xji;
    const cpx w = wN[k];
    const double cr = dr * w.x + di * w.y, ci = di * w.x - dr * w.y;     // conj(w) D
    const double tr = -ci, ti = cr;                                       // i * (.)
    dst.st(pk, sr + tr, si + ti);
    dst.st(pj, sr - tr, -(si - ti));
  }
}

// ---- type IV (ND / DN): no split -- a pre-twiddle before and a post-twiddle after the complex FFT, both directions.
// wQ[m] = e^{-i pi (4m+1)/(4N)}, wN[k] = e^{-i pi k/N} (M entries each for these kinds).
// iv_pre (natural slot order): slot m <- z_m wQ[m]; BWD: the spectrum sits digit-reversed (slot pos[m]), the row pair is
// (Y_{2m}, Y_{N-1-2m}) (DN: swapped), and the DIT passes that follow compute the +i transform: feed conj(v).
template <bool FWD, class ACC>
FB_HD void iv_pre(int M, int kind, const cpx* wQ, const int* pos, int worker, int nworkers, const ACC& acc) {
  for (int m = worker; m < M; m += nworkers) {
    const int sl = FWD ? m : pos[m];
    double zr, zi;
    acc.ld(sl, zr, zi);
    if (!FWD && kind == KIND_DN) { const double t = zr; zr = zi; zi = t; }
    const cpx w = wQ[m];
    const double vr = zr * w.x - zi * w.y, vi = zr * w.y + zi * w.x;
    acc.st(sl, vr, FWD ? vi : -vi);
  }
}
// iv_post: FWD: mode k sits at slot pos[k] (DIF output); BWD: at slot k, conjugated (DIT of the conjugate).
// (a, b) = (2 Re w, -2 Im w): FWD rows (Y_{2k}, Y_{N-1-2k}); BWD elements (x_{2k}, x_{N-1-2k}); DN: second sign +, and the
// physical pair is stored swapped (the tile's part 0 is element N-1-2k of a reversed line).
template <bool FWD, class ACC>
FB_HD void iv_post(int M, int kind, const cpx* wN, const int* pos, int worker, int nworkers, const ACC& acc) {
  for (int k = worker; k < M; k += nworkers) {
    const int sl = FWD ? pos[k] : k;
    double vr, vi;
    acc.ld(sl, vr, vi);
    if (!FWD) vi = -vi;
    const cpx w = wN[k];
    const double a = 2.0 * (vr * w.x - vi * w.y), b = -2.0 * (vr * w.y + vi * w.x);
    if (kind == KIND_DN) { if (FWD) acc.st(sl, a, -b); else acc.st(sl, -b, a); }
    else acc.st(sl, a, b);
  }
}

template <int TB, bool ROT>
FB_HD void split_fwd(double* tile, const LinePlan& P, int lane, int worker, int nworkers) {
  const TileAcc<TB, ROT> acc{tile, P.M, lane};
  split_core(P.M, P.kind, P.wN, P.wQ, P.pos, worker, nworkers, acc, acc);
}

template <int TB, bool ROT>
FB_HD void merge_bwd(double* tile, const LinePlan& P, int lane, int worker, int nworkers) {
  const TileAcc<TB, ROT> acc{tile, P.M, lane};
  merge_core(P.M, P.kind, P.wN, P.wQ, P.pos, worker, nworkers, acc, acc);
}

// ---------------------------------------------------------------------------------------------
// power-of-two lengths: the radix schedule is a compile-time function of M (identical to the one
// make_line_plan builds at run time: 8,8,...,then 4 or 2), so strides/twiddle steps fold to constants.
FB_CX int p2_first_radix(int Lc) { return (Lc % 8 == 0) ? 8 : (Lc % 4 == 0) ? 4 : 2; }
FB_CX int p2_npass(int M) { int n = 0; while (M > 1) { M /= p2_first_radix(M); ++n; } return n; }
FB_CX int p2_lc(int M, int q) { while (q-- > 0) M /= p2_first_radix(M); return M; }   // sub-length before pass q

template <int M, int Q, bool FWD, class LD, class ST>
FB_HD void p2_pass(const cpx* wM, int worker, int nworkers, const LD& src, const ST& dst) {
  constexpr int Lc = p2_lc(M, Q), R = p2_first_radix(Lc), s = Lc / R;
  pass_core<R, FWD>(M, s, wM, worker, nworkers, src, dst);
}

}  // namespace fb
