#include <cstdio>
// CPU emulation of the tile-FFT core (flutas_b200/csrc/tile_fft.cuh) for the "not gpu" tests.
// TEST INFRASTRUCTURE: runs the kernels' per-thread phase functions in a serial loop over
// (lane, worker) with the phase boundaries where the CUDA kernels have __syncthreads().  It lets
// the index logic (Makhoul permutation, rotation swizzle, digit reversal, split/merge, mode map) be
// checked against the oracle without a GPU.  Nothing in the product links this file.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../flutas_b200/csrc/line_plan.h"

using namespace fb;

template <int TB, bool ROT>
static void run(const HostLinePlan& hp, int nworkers, int fwd, const double* in, double* out, double scale) {
  LinePlan P;
  P.N = hp.N; P.M = hp.M; P.kind = hp.kind; P.npass = (int)hp.radix.size();
  for (int q = 0; q < P.npass; ++q) { P.radix[q] = hp.radix[q]; P.sub[q] = hp.sub[q]; }
  P.wM = hp.wM.data(); P.wN = hp.wN.data(); P.wQ = hp.wQ.data(); P.pos = hp.pos.data();
  const int N = hp.N, M = hp.M;
  std::vector<double> tile((size_t)N * TB, 0.0);
  // load phase
  for (int L = 0; L < TB; ++L)
    for (int e = 0; e < N; ++e) {
      int m, part; double sgn = 1.0;
      if (fwd) elem_to_slot(hp.kind, N, e, m, part, sgn);
      else { part = (e >= M); m = e - part * M; }
      tile[taddr<TB, ROT>(m, part, M, L)] = sgn * in[(size_t)L * N + e];
    }
  auto phase = [&](auto&& fn) {
    for (int w = 0; w < nworkers; ++w)
      for (int lane = 0; lane < TB; ++lane) fn(lane, w);
  };
  const bool iv = kind_is_iv(P.kind);                     // same phase order as tile_transform (kernels.cuh)
  auto acc = [&](int lane) { return TileAcc<TB, ROT>{tile.data(), P.M, lane}; };
  if (fwd) {
    if (iv) phase([&](int lane, int w) { iv_pre<true>(P.M, P.kind, P.wQ, P.pos, w, nworkers, acc(lane)); });
    for (int q = 0; q < P.npass; ++q)
      phase([&](int lane, int w) { fft_pass<TB, ROT, true>(tile.data(), P, P.wM, q, lane, w, nworkers); });
    if (iv) phase([&](int lane, int w) { iv_post<true>(P.M, P.kind, P.wN, P.pos, w, nworkers, acc(lane)); });
    else phase([&](int lane, int w) { split_fwd<TB, ROT>(tile.data(), P, lane, w, nworkers); });
  } else {
    if (iv) phase([&](int lane, int w) { iv_pre<false>(P.M, P.kind, P.wQ, P.pos, w, nworkers, acc(lane)); });
    else phase([&](int lane, int w) { merge_bwd<TB, ROT>(tile.data(), P, lane, w, nworkers); });
    for (int q = P.npass - 1; q >= 0; --q)
      phase([&](int lane, int w) { fft_pass<TB, ROT, false>(tile.data(), P, P.wM, q, lane, w, nworkers); });
    if (iv) phase([&](int lane, int w) { iv_post<false>(P.M, P.kind, P.wN, P.pos, w, nworkers, acc(lane)); });
  }
  // store phase
  for (int L = 0; L < TB; ++L)
    for (int e = 0; e < N; ++e) {
      int m, part; double sgn = 1.0;
      if (!fwd) elem_to_slot(hp.kind, N, e, m, part, sgn);
      else { part = (e >= M); m = e - part * M; }
      out[(size_t)L * N + e] = sgn * scale * tile[taddr<TB, ROT>(m, part, M, L)];
    }
}

extern "C" int emul_line_transform(int N, int kind, int TB, int rot, int nworkers, int fwd,
                                   const double* in, double* out, double scale) {
  HostLinePlan hp = make_line_plan(N, kind);
  if (!hp.ok) return 1;
#define CASE(tb) \
  if (TB == tb) { if (rot) run<tb, true>(hp, nworkers, fwd, in, out, scale); else run<tb, false>(hp, nworkers, fwd, in, out, scale); return 0; }
  CASE(4) CASE(8) CASE(16)
#undef CASE
  return 2;
}

extern "C" int emul_mode_index(int N, int kind, int* mode) {
  HostLinePlan hp = make_line_plan(N, kind);
  if (!hp.ok) return 1;
  std::memcpy(mode, hp.mode.data(), sizeof(int) * (size_t)N);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// on-chip Thomas (flutas_b200/csrc/thomas_tile.cuh): same phases as thomas_tile_kernel, serial over threads
#include "../../flutas_b200/csrc/thomas_tile.cuh"

template <int L, int TI>
static void thomas_emul(long ncol, ThomasArgs T, const double* lam, double* W) {
  using TT = ThomasTile<L, TI>;
  const int nz = T.nz, S = T.S;
  const long nblk = (ncol + TI - 1) / TI;
  std::vector<double> smem(TT::smem_doubles(nz));
  std::vector<double> zreg((size_t)S * TI * (L > 1 ? L - 1 : 1));
  for (long blk = 0; blk < nblk; ++blk) {
    double* tile = smem.data();
    double* exa = tile + (size_t)TT::tile_rows(nz) * TI;
    double* exb = exa + 4 * (size_t)S * TI;
    const long col0 = blk * TI;
    auto live = [&](int lane) { return col0 + lane < ncol; };
    auto lamof = [&](int lane) { return live(lane) ? lam[col0 + lane] : -1.0; };
    auto zof = [&](int lane, int s) { return zreg.data() + ((size_t)s * TI + lane) * (L > 1 ? L - 1 : 1); };
    for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane)
      for (int k = s; k < nz; k += S) tile[TT::prow(k) * TI + lane] = live(lane) ? W[col0 + lane + (long)k * ncol] : 0.0;
    if (L > 1)
      for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane)
        TT::local_sweeps(tile, exa, T, lamof(lane), lane, s, zof(lane, s));
    // reduced rows: computed by every thread from exa, then written over exa (kernel keeps them in registers)
    std::vector<double> red(4 * (size_t)S * TI);
    for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane) {
      const bool pin = T.singular && live(lane) && lamof(lane) == 0.0;
      TT::reduced_row(tile, exa, red.data(), T, lamof(lane), lane, s, pin);
    }
    std::memcpy(exa, red.data(), sizeof(double) * red.size());
    double* src = exa; double* dst = exb;
    const int hmax = T.periodic ? S / 2 : S;
    for (int h = 1; h < hmax; h *= 2) {
      for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane) TT::pcr_step(src, dst, T, lane, s, h);
      double* t = src; src = dst; dst = t;
    }
    for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane) TT::pcr_finish(src, dst, T, lane, s);
    for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane) TT::substitute(tile, dst, T, lane, s, zof(lane, s));
    for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane)
      if (live(lane)) for (int k = s; k < nz; k += S) W[col0 + lane + (long)k * ncol] = tile[TT::prow(k) * TI + lane];
  }
}

extern "C" int emul_thomas_tile(int L, int nz, long ncol, int periodic, int singular, const double* a, const double* b,
                                const double* c, const double* lam, double* W) {
  if (nz % L) return 1;
  std::vector<double> az(a, a + nz), cz(c, c + nz);
  if (!periodic) { az[0] = 0.0; cz[nz - 1] = 0.0; }
  ThomasArgs T;
  T.nz = nz; T.S = nz / L; T.periodic = periodic; T.singular = singular; T.az = az.data(); T.bz = b; T.cz = cz.data();
  T.padded = 0;
  if (periodic && (T.S & (T.S - 1))) return 2;
  switch (L) {
    case 1: thomas_emul<1, 8>(ncol, T, lam, W); break;
    case 2: thomas_emul<2, 8>(ncol, T, lam, W); break;
    case 4: thomas_emul<4, 8>(ncol, T, lam, W); break;
    case 8: thomas_emul<8, 8>(ncol, T, lam, W); break;
    case 16: thomas_emul<16, 8>(ncol, T, lam, W); break;
    case 32: thomas_emul<32, 8>(ncol, T, lam, W); break;
    default: return 3;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// register Thomas (flutas_b200/csrc/thomas_reg.cuh): same phases as thomas_reg_kernel, serial over threads
#include "../../flutas_b200/csrc/thomas_reg.cuh"

template <int L, int TI, bool UNI>
static void thomas_reg_emul(long ncol, ThomasArgs T, const double* lam, double* W) {
  using TR = ThomasReg<L, TI>;
  using CF = typename std::conditional<UNI, CoefUniform<L>, CoefTable<L>>::type;
  const int nz = T.nz, S = T.S, st = S * TI;
  const long ntiles = (ncol + TI - 1) / TI;
  // padded coefficient rows, as the kernel stages them
  const int tr = TR::tile_rows(nz);
  std::vector<double> coef(3 * (size_t)tr, 0.0);
  for (int k = 0; k < nz; ++k) { const int r = TR::prow(k); coef[r] = T.az[k]; coef[tr + r] = T.bz[k]; coef[2 * tr + r] = T.cz[k]; }
  T.az = coef.data(); T.bz = coef.data() + tr; T.cz = coef.data() + 2 * tr; T.padded = 1;
  std::vector<double> ex(6 * (size_t)st), pa(3 * (size_t)st), pb(3 * (size_t)st), X(st);
  std::vector<SegRegs<L>> regs(st);
  std::vector<double> v((size_t)st * L);
  for (long tile = 0; tile < ntiles; ++tile) {
    auto colof = [&](int lane) { return tile * TI + lane; };
    auto live = [&](int lane) { return colof(lane) < ncol; };
    auto lamof = [&](int lane) { return live(lane) ? lam[colof(lane)] : -1.0; };
    auto vof = [&](int lane, int s) { return v.data() + ((size_t)s * TI + lane) * L; };
#define ALLT for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane)
    ALLT { const long col = live(lane) ? colof(lane) : ncol - 1; for (int l = 0; l < L; ++l) vof(lane, s)[l] = W[col + (long)(s * L + l) * ncol]; }
    ALLT TR::phase1(vof(lane, s), T, CF(T, s), lamof(lane), lane, s, regs[s * TI + lane], ex.data());
    ALLT { const bool pin = T.singular && live(lane) && lamof(lane) == 0.0;
           TR::reduced_row(vof(lane, s)[L - 1], ex.data(), pa.data(), T, CF(T, s), lamof(lane), lane, s, pin); }
    double* src = pa.data(); double* dst = pb.data();
    const int hmax = T.periodic ? S / 2 : S;
    for (int h = 1; h < hmax; h *= 2) {
      bool any = false;
      ALLT any = TR::pcr_step(src, dst, T, lane, s, h) || any;
      double* t = src; src = dst; dst = t;
      if (!any) break;
    }
    ALLT TR::pcr_finish(src, X.data(), T, lane, s);
    ALLT TR::phase3(vof(lane, s), X.data(), T, lane, s, regs[s * TI + lane]);
    ALLT if (live(lane)) for (int l = 0; l < L; ++l) W[colof(lane) + (long)(s * L + l) * ncol] = vof(lane, s)[l];
#undef ALLT
  }
}

extern "C" int emul_thomas_reg(int L, int nz, long ncol, int periodic, int singular, const double* a, const double* b,
                               const double* c, const double* lam, double* W, int allow_uniform) {
  if (nz % L) return 1;
  std::vector<double> az(a, a + nz), cz(c, c + nz);
  if (!periodic) { az[0] = 0.0; cz[nz - 1] = 0.0; }
  ThomasArgs T;
  T.nz = nz; T.S = nz / L; T.periodic = periodic; T.singular = singular; T.az = az.data(); T.bz = b; T.cz = cz.data();
  T.padded = 0;
  T.uniform = 0;
  if (T.S < 2) return 2;
  if (periodic && (T.S & (T.S - 1))) return 2;
  const bool uni = allow_uniform && thomas_detect_uniform(nz, a, b, c, periodic != 0, T);
  if (allow_uniform == 2 && !uni) return 4;                  // the caller expected the uniform path
#define RUN(LL) if (uni) thomas_reg_emul<LL, 8, true>(ncol, T, lam, W); else thomas_reg_emul<LL, 8, false>(ncol, T, lam, W); break;
  switch (L) {
    case 2: RUN(2)
    case 4: RUN(4)
    case 8: RUN(8)
    case 16: RUN(16)
    default: return 3;
  }
#undef RUN
  return 0;
}

// shared-LU kernel for exactly uniform grids (flutas_b200/csrc/thomas_uni.cuh): same phases as thomas_uni_kernel
#include "../../flutas_b200/csrc/thomas_uni.cuh"

template <int L, int TI>
static void thomas_uni_emul(long ncol, ThomasArgs T, const double* lam, double* W) {
  using TU = ThomasUni<L, TI>;
  using TR = ThomasReg<L, TI>;
  const int S = T.S, st = S * TI;
  const long ntiles = (ncol + TI - 1) / TI;
  std::vector<double> tab(TU::tab_doubles()), ex(6 * (size_t)st), pa(3 * (size_t)st), pb(3 * (size_t)st), X(st);
  std::vector<double> v((size_t)st * L);
  for (long tile = 0; tile < ntiles; ++tile) {
    auto colof = [&](int lane) { return tile * TI + lane; };
    auto live = [&](int lane) { return colof(lane) < ncol; };
    auto lamof = [&](int lane) { return live(lane) ? lam[colof(lane)] : -1.0; };
    auto vof = [&](int lane, int s) { return v.data() + ((size_t)s * TI + lane) * L; };
#define ALLT for (int s = 0; s < S; ++s) for (int lane = 0; lane < TI; ++lane)
    std::fill(tab.begin(), tab.end(), std::nan(""));
    for (int var = 0; var < 2; ++var) for (int lane = 0; lane < TI; ++lane) TU::build(tab.data(), T, lamof(lane), lane, var);
    ALLT { const long col = live(lane) ? colof(lane) : ncol - 1; for (int l = 0; l < L; ++l) vof(lane, s)[l] = W[col + (long)(s * L + l) * ncol]; }
    ALLT TU::phase1(vof(lane, s), tab.data(), T, lane, s, ex.data());
    ALLT { const bool pin = T.singular && live(lane) && lamof(lane) == 0.0;
           TR::reduced_row(vof(lane, s)[L - 1], ex.data(), pa.data(), T, CoefUniform<L>(T, s), lamof(lane), lane, s, pin); }
    double* src = pa.data(); double* dst = pb.data();
    const int hmax = T.periodic ? S / 2 : S;
    for (int h = 1; h < hmax; h *= 2) {
      bool any = false;
      ALLT any = TR::pcr_step(src, dst, T, lane, s, h) || any;
      double* t = src; src = dst; dst = t;
      if (!any) break;
    }
    ALLT TR::pcr_finish(src, X.data(), T, lane, s);
    ALLT TU::phase3(vof(lane, s), X.data(), tab.data(), T, lane, s);
    ALLT if (live(lane)) for (int l = 0; l < L; ++l) W[colof(lane) + (long)(s * L + l) * ncol] = vof(lane, s)[l];
#undef ALLT
  }
}

// returns the segment length used (> 0), 0 if the grid is not exactly uniform / nz not served
extern "C" int emul_thomas_uni(int nz, long ncol, int periodic, int singular, const double* a, const double* b,
                               const double* c, const double* lam, double* W) {
  ThomasArgs T;
  T.nz = nz; T.periodic = periodic; T.singular = singular; T.az = T.bz = T.cz = nullptr; T.padded = 0;
  if (!thomas_detect_uniform(nz, a, b, c, periodic != 0, T)) return 0;
  int L = 0;
  if (!thomas_uni_pick(nz, periodic != 0, &L)) return 0;
  T.S = nz / L;
  switch (L) {
    case 4: thomas_uni_emul<4, 16>(ncol, T, lam, W); break;
    case 8: thomas_uni_emul<8, 16>(ncol, T, lam, W); break;
    case 16: thomas_uni_emul<16, 16>(ncol, T, lam, W); break;
    default: thomas_uni_emul<32, 16>(ncol, T, lam, W); break;
  }
  return L;
}

extern "C" int emul_thomas_reg_pick(int nz, int periodic) {
  int L = 0;
  return thomas_reg_pick(nz, periodic != 0, &L) ? L : 0;
}

// ---------------------------------------------------------------------------------------------
// register-resident transforms (flutas_b200/csrc/reg_fft.cuh): the per-thread functions the kernels call, run
// serially over the T threads of one line with the phase boundaries where the kernels synchronise.
template <bool SWZ>
struct EmulXB {                          // exchange buffer of one line: padded like the y tiles' (they add lanes) or
  static constexpr bool XOR = SWZ;       // XOR-swizzled like the x lines'
  double* re; double* im;
  static constexpr int off(int c) { return SWZ ? rf_swzoff(c) : rf_padoff(c); }
  int base(int pos) const { return SWZ ? rf_swz(pos) : rf_pad(pos); }
  static int at(int b, int coff) { return SWZ ? (b ^ coff) : (b + coff); }
  static constexpr int offsub(int a, int b) { return SWZ ? (rf_swz(a) ^ rf_swz(b)) : (rf_padoff(a) - rf_padoff(b)); }
  int addr(int b, int coff) const { return at(b, coff); }
  void st(int b, int coff, double r, double i) const { re[at(b, coff)] = r; im[at(b, coff)] = i; }
  void ld(int b, int coff, double& r, double& i) const { r = re[at(b, coff)]; i = im[at(b, coff)]; }
};

template <int M, int RR = 16, bool SWZ = false>
static void reg_line_emul1(const HostRegPlan& hp, int fwd, const double* in, double* out, double scale, bool pair = false) {
  using S = RegSched<M, RR>;
  constexpr int T = S::T, R = S::R, N = 2 * M;
  // poisoned padded buffer: a wrong padded address reads NaN (or clobbers a slot that is read later)
  std::vector<double> bre(M + M / 16 + 2, std::nan("")), bim(M + M / 16 + 2, std::nan(""));
  EmulXB<SWZ> xb{bre.data(), bim.data()};
  std::vector<double> re((size_t)T * R), im((size_t)T * R);
  auto RE = [&](int j) { return re.data() + (size_t)j * R; };
  auto IM = [&](int j) { return im.data() + (size_t)j * R; };
  const cpx* tw[RF_MAXPASS] = {nullptr, RR == 16 ? hp.tw[1].data() : hp.tw8[1].data(), RR == 16 ? hp.tw[2].data() : hp.tw8[2].data()};
  auto passes = [&](auto sign) {
    constexpr int SIGN = decltype(sign)::value;
    for (int j = 0; j < T; ++j) reg_pass<S, 0, SIGN>(RE(j), IM(j), j, tw[0], xb);
    if constexpr (S::NP > 1) {
      for (int j = 0; j < T; ++j) reg_gather<S>(RE(j), IM(j), j, xb);
      for (int j = 0; j < T; ++j) reg_pass<S, 1, SIGN>(RE(j), IM(j), j, tw[1], xb);
    }
    if constexpr (S::NP > 2) {
      for (int j = 0; j < T; ++j) reg_gather<S>(RE(j), IM(j), j, xb);
      for (int j = 0; j < T; ++j) reg_pass<S, 2, SIGN>(RE(j), IM(j), j, tw[2], xb);
    }
  };
  if (fwd) {
    for (int j = 0; j < T; ++j)
      for (int u = 0; u < R; ++u) {
        int e0, e1; double s0, s1;
        reg_phys_slots(hp.kind, N, j + T * u, e0, e1, s0, s1);
        if (kind_is_makhoul(hp.kind) && M >= 16) {         // the kernels' closed form of the same rows
          using MR = MkRows<N, RR>;
          const int b = MR::upper(u) ? MR::base_hi(j) : MR::base_lo(j);
          if (b + MR::off0(u) != e0 || b + MR::off1(u) != e1) std::abort();
          const double sg = (hp.kind == KIND_DD && MR::upper(u)) ? -1.0 : 1.0;
          if (sg != s0 || sg != s1) std::abort();
        }
        RE(j)[u] = s0 * in[e0]; IM(j)[u] = s1 * in[e1];
      }
    const bool iv = kind_is_iv(hp.kind), dn = (hp.kind == KIND_DN);
    if (iv) for (int j = 0; j < T; ++j) reg_iv_pre<S, false>(RE(j), IM(j), j, hp.wQ.data());
    if (!iv && pair && reg_has_pair_pass<S>()) {          // small-radix last pass on symmetric butterfly pairs + split in registers
      if constexpr (reg_has_pair_pass<S>()) {
        for (int j = 0; j < T; ++j) reg_pass<S, 0, -1>(RE(j), IM(j), j, tw[0], xb);
        if constexpr (S::NP > 2) {
          for (int j = 0; j < T; ++j) reg_gather<S>(RE(j), IM(j), j, xb);
          for (int j = 0; j < T; ++j) reg_pass<S, 1, -1>(RE(j), IM(j), j, tw[1], xb);
        }
        std::vector<int> seen(M, 0);
        auto put = [&](int k, double xr, double xi) { out[2 * k] = scale * xr; out[2 * k + 1] = scale * xi; seen[k]++; };
        for (int j = 0; j < T; ++j) {
          if (hp.kind == KIND_PP) reg_pair_pass_split<S, false>(j, tw[S::NP - 1], hp.wN.data(), hp.wQ.data(), xb, put);
          else reg_pair_pass_split<S, true>(j, tw[S::NP - 1], hp.wN.data(), hp.wQ.data(), xb, put);
        }
        for (int k = 0; k < M; ++k) if (seen[k] != 1) std::abort();      // every mode exactly once
      }
      return;
    }
    passes(std::integral_constant<int, -1>{});
    if (iv) {
      for (int j = 0; j < T; ++j) reg_iv_post<S, true>(RE(j), IM(j), j, hp.wN.data(), dn);
    } else {
      for (int j = 0; j < T; ++j) reg_scatter_modes<S>(RE(j), IM(j), j, xb);
      for (int j = 0; j < T; ++j) {
        if (hp.kind == KIND_PP) reg_split<S, false>(RE(j), IM(j), j, hp.wN.data(), hp.wQ.data(), xb);
        else reg_split<S, true>(RE(j), IM(j), j, hp.wN.data(), hp.wQ.data(), xb);
      }
    }
    for (int j = 0; j < T; ++j)
      for (int u = 0; u < R; ++u) { const int k = j + T * u; out[2 * k] = scale * RE(j)[u]; out[2 * k + 1] = scale * IM(j)[u]; }
  } else {
    for (int j = 0; j < T; ++j)
      for (int u = 0; u < R; ++u) {
        const int k = j + T * u;
        const bool dn = (hp.kind == KIND_DN);
        RE(j)[u] = dn ? in[2 * k + 1] : in[2 * k]; IM(j)[u] = dn ? in[2 * k] : in[2 * k + 1];
      }
    const bool iv = kind_is_iv(hp.kind);
    bool transposed = false;
    if constexpr (reg_has_pair_pass<S>()) {
      if (!iv && pair) {                                   // merge in registers + transposed schedule
        transposed = true;
        std::vector<int> seen(M, 0);
        auto get = [&](int k, double& xr, double& xi) { xr = in[2 * k]; xi = in[2 * k + 1]; seen[k]++; };
        for (int j = 0; j < T; ++j) {
          if (hp.kind == KIND_PP) reg_pair_merge_pass<S, false>(j, tw[S::NP - 1], hp.wN.data(), hp.wQ.data(), xb, get);
          else reg_pair_merge_pass<S, true>(j, tw[S::NP - 1], hp.wN.data(), hp.wQ.data(), xb, get);
        }
        for (int k = 0; k < M; ++k) if (seen[k] != 1) std::abort();
        if constexpr (S::NP > 2) {
          for (int j = 0; j < T; ++j) reg_pass_T_load<S, 1>(RE(j), IM(j), j, xb);
          for (int j = 0; j < T; ++j) reg_pass_T_finish<S, 1, +1>(RE(j), IM(j), j, tw[1], xb);
        }
        for (int j = 0; j < T; ++j) reg_pass_T_load<S, 0>(RE(j), IM(j), j, xb);
        for (int j = 0; j < T; ++j) reg_pass_T_finish<S, 0, +1>(RE(j), IM(j), j, tw[0], xb);
      }
    }
    if (transposed) {
    } else if (iv) {
      for (int j = 0; j < T; ++j) reg_iv_pre<S, true>(RE(j), IM(j), j, hp.wQ.data());
    } else {
      for (int j = 0; j < T; ++j) reg_scatter_modes<S>(RE(j), IM(j), j, xb);
      for (int j = 0; j < T; ++j) {
        if (hp.kind == KIND_PP) reg_merge<S, false>(RE(j), IM(j), j, hp.wN.data(), hp.wQ.data(), xb);
        else reg_merge<S, true>(RE(j), IM(j), j, hp.wN.data(), hp.wQ.data(), xb);
      }
    }
    if (!transposed) passes(std::integral_constant<int, +1>{});
    if (iv) for (int j = 0; j < T; ++j) reg_iv_post<S, false>(RE(j), IM(j), j, hp.wN.data(), hp.kind == KIND_DN);
    for (int j = 0; j < T; ++j)
      for (int u = 0; u < R; ++u) {
        int e0, e1; double s0, s1;
        reg_phys_slots(hp.kind, N, j + T * u, e0, e1, s0, s1);
        out[e0] = s0 * scale * RE(j)[u]; out[e1] = s1 * scale * IM(j)[u];
      }
  }
}

static bool g_pair = false;           // emul_reg_pair_mode: forward lines through reg_pair_pass_split where the schedule has it
extern "C" int emul_reg_pair_mode(int on) { g_pair = on != 0; return 0; }
// both exchange-buffer addressings (padded: y tiles; XOR-swizzled: x lines) must give the same bits
template <int M, int RR = 16>
static void reg_line_emul(const HostRegPlan& hp, int fwd, const double* in, double* out, double scale) {
  reg_line_emul1<M, RR, true>(hp, fwd, in, out, scale, g_pair);
  std::vector<double> alt((size_t)2 * M);
  reg_line_emul1<M, RR, false>(hp, fwd, in, alt.data(), scale, g_pair);
  if (std::memcmp(alt.data(), out, sizeof(double) * 2 * M) != 0) {
    std::fprintf(stderr, "exchange-buffer addressings disagree: M=%d RR=%d kind=%d fwd=%d\n", M, RR, hp.kind, fwd);
    std::abort();
  }
}

extern "C" int emul_reg_line_transform(int N, int kind, int fwd, const double* in, double* out, double scale) {
  HostRegPlan hp = make_reg_plan(N, kind);
  if (!hp.ok) return 1;
  switch (hp.M) {
    case 16: reg_line_emul<16>(hp, fwd, in, out, scale); break;
    case 32: reg_line_emul<32>(hp, fwd, in, out, scale); break;
    case 64: reg_line_emul<64>(hp, fwd, in, out, scale); break;
    case 128: reg_line_emul<128>(hp, fwd, in, out, scale); break;
    case 256: reg_line_emul<256>(hp, fwd, in, out, scale); break;
    case 512: reg_line_emul<512>(hp, fwd, in, out, scale); break;
    case 1024: reg_line_emul<1024>(hp, fwd, in, out, scale); break;
    default: return 2;
  }
  return 0;
}

// the 8-values-per-thread schedule (N = 1024)
extern "C" int emul_reg_line_transform8(int N, int kind, int fwd, const double* in, double* out, double scale) {
  HostRegPlan hp = make_reg_plan(N, kind);
  if (!hp.ok || hp.M != 512) return 1;
  reg_line_emul<512, 8>(hp, fwd, in, out, scale);
  return 0;
}

extern "C" int emul_reg_mode_index(int N, int kind, int* mode) {
  HostRegPlan hp = make_reg_plan(N, kind);
  if (!hp.ok) return 1;
  std::memcpy(mode, hp.mode.data(), sizeof(int) * (size_t)N);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// reference-order solve of the ill-conditioned columns (flutas_b200/csrc/thomas_ref.cuh): same selection rule as
// capi.cu's ensure_ref, same per-lane phase functions as ref_solve_kernel run serially over the 32 lanes of a warp.
// Overwrites the selected columns of W (ncol x nz); returns the number of selected columns.
#include "../../flutas_b200/csrc/thomas_ref.cuh"

extern "C" int emul_thomas_ref(int nz, long ncol, int periodic, int singular, const double* a, const double* b,
                               const double* c, const double* lam, double* W, double tol) {
  double amax = 0.0;
  for (int k = 0; k < nz; ++k) amax = std::fmax(amax, std::fmax(std::fabs(a[k]), std::fabs(c[k])));
  const double thr = 4.0 * amax * tol;
  std::vector<int> sel;
  for (long q = 0; q < ncol; ++q) if (std::fabs(lam[q]) < thr) sel.push_back((int)q);
  const int ns = (int)sel.size();
  if (!ns) return 0;
  std::vector<double> sl(ns), z((size_t)ns * nz), d((size_t)ns * nz), piv(ns), p2((size_t)ns * nz), den(ns);
  std::vector<unsigned char> pin(ns);
  for (int q = 0; q < ns; ++q) { sl[q] = lam[sel[q]]; pin[q] = (singular && sl[q] == 0.0) ? 1 : 0; }
  RefTables R;
  R.nsel = ns; R.nz = nz; R.m = periodic ? nz - 1 : nz; R.periodic = periodic;
  R.a = a; R.b = b; R.c = c; R.col = sel.data(); R.lam = sl.data(); R.pin = pin.data();
  R.z = z.data(); R.d = d.data(); R.piv = piv.data(); R.p2 = p2.data(); R.den = den.data();
  for (int q = 0; q < ns; ++q) ref_factor(R, q);
  const RefShape S(nz);
  std::vector<double> p(S.doubles(), std::nan("")), zs(S.doubles(), std::nan("")), ds(S.doubles(), std::nan(""));
  double PA[32], PB[32], IN[32];
  const int n = nz, m = R.m;
  for (int q = 0; q < ns; ++q) {
    const long col = sel[q];
    for (int l = 0; l < n; ++l) {
      const int o = S.at(l);
      p[o] = W[col + ncol * (long)l];
      if (l < m - 1) { zs[o] = z[(size_t)q * n + l]; ds[o] = d[(size_t)q * n + l]; }
    }
    for (int lane = 0; lane < 32; ++lane) ref_fwd_local(R, S, p.data(), zs.data(), lane, PA, PB);
    ref_chain_up(PA, PB, IN, 0.0);
    for (int lane = 0; lane < 32; ++lane) ref_fwd_final(R, S, p.data(), zs.data(), lane, IN[lane]);
    const double xm = ref_last_row(R, S, p.data(), q, pin[q] && !periodic);
    for (int lane = 0; lane < 32; ++lane) ref_bwd_local(R, S, p.data(), ds.data(), lane, PA, PB);
    ref_chain_down(PA, PB, IN, xm);
    for (int lane = 0; lane < 32; ++lane) ref_bwd_final(R, S, p.data(), ds.data(), lane, IN[lane]);
    if (periodic) {
      const double pn = ref_closure(R, S, p.data(), q, pin[q] != 0);
      for (int l = 0; l < m; ++l) W[col + ncol * (long)l] = FB_XADD(p[S.at(l)], FB_XMUL(p2[(size_t)q * n + l], pn));
      W[col + ncol * (long)(n - 1)] = pn;
    } else {
      for (int l = 0; l < n; ++l) W[col + ncol * (long)l] = p[S.at(l)];
    }
  }
  return ns;
}

// ---------------------------------------------------------------------------------------------
// distributed z solve (flutas_b200/csrc/dz.cuh): G ranks each own nz/G consecutive levels of every column.  Same steps
// as capi.cu's dz_setup / dz_solve_z with the rank-local solves done by the shared-LU kernel's phase functions
// (thomas_uni_emul<16,16>, exactly what thomas_uni_local_run launches) and the interface systems by dz_interface_solve.
#include "../../flutas_b200/csrc/dz.cuh"

// general == 1: the local block through the general kernel's phase functions (coefficient tables, thomas_reg_local_run)
static int g_dz_general = 0;
static const double *g_dz_a = nullptr, *g_dz_b = nullptr, *g_dz_c = nullptr;     // local rows of a, b, c for the general path
static bool dz_local_solve(int n3l, long ncol, const ThomasArgs& uni, int singular, const double* lam, double* slab) {
  if (g_dz_general) return emul_thomas_reg(16, n3l, ncol, 0, singular, g_dz_a, g_dz_b, g_dz_c, lam, slab, 0) == 0;
  ThomasArgs T = uni;
  T.nz = n3l; T.S = n3l / 16; T.periodic = 0; T.singular = singular; T.az = T.bz = T.cz = nullptr; T.padded = 0;
  if (n3l % 16 || T.S < 2 || T.S > 32) return false;
  thomas_uni_emul<16, 16>(ncol, T, lam, slab);
  return true;
}

extern "C" int emul_dz(int nz, int G, long ncol, int periodic, int singular, const double* a, const double* b, const double* c,
                       const double* lam, double* W) {
  if (G < 2 || G > FB_DZ_MAXG || nz % G) return 1;
  const int n3l = nz / G;
  std::vector<ThomasArgs> uni(G);
  std::vector<double> ca(G), cc(G);
  std::vector<double> PF((size_t)G * ncol), PL(PF), QF(PF), QL(PF), scratch((size_t)ncol * n3l);
  for (int g = 0; g < G; ++g) {
    const int k0 = g * n3l;
    if (!g_dz_general && !thomas_detect_uniform(n3l, a + k0, b + k0, c + k0, false, uni[g])) return 2;
    g_dz_a = a + k0; g_dz_b = b + k0; g_dz_c = c + k0;
    ca[g] = (periodic || g > 0) ? a[k0] : 0.0;
    cc[g] = (periodic || g < G - 1) ? c[k0 + n3l - 1] : 0.0;
    const int sing = (singular && g == G - 1) ? 1 : 0;
    for (int pass = 0; pass < 2; ++pass) {
      std::fill(scratch.begin(), scratch.end(), 0.0);
      for (long q = 0; q < ncol; ++q) scratch[(pass == 0 ? 0 : (size_t)ncol * (n3l - 1)) + q] = pass == 0 ? ca[g] : cc[g];
      if (!dz_local_solve(n3l, ncol, uni[g], sing, lam, scratch.data())) return 3;
      for (long q = 0; q < ncol; ++q) {
        (pass == 0 ? PF : QF)[(size_t)g * ncol + q] = scratch[q];
        (pass == 0 ? PL : QL)[(size_t)g * ncol + q] = scratch[(size_t)ncol * (n3l - 1) + q];
      }
    }
  }
  // pass 1 on every rank
  for (int g = 0; g < G; ++g) {
    g_dz_a = a + g * n3l; g_dz_b = b + g * n3l; g_dz_c = c + g * n3l;
    if (!dz_local_solve(n3l, ncol, uni[g], (singular && g == G - 1) ? 1 : 0, lam, W + (size_t)ncol * g * n3l)) return 3;
  }
  // interface systems
  std::vector<double> XP((size_t)G * ncol), XN((size_t)G * ncol);
  for (long q = 0; q < ncol; ++q) {
    double pF[FB_DZ_MAXG], pL[FB_DZ_MAXG], qF[FB_DZ_MAXG], qL[FB_DZ_MAXG], yF[FB_DZ_MAXG], yL[FB_DZ_MAXG], u[2 * FB_DZ_MAXG];
    for (int g = 0; g < G; ++g) {
      pF[g] = PF[(size_t)g * ncol + q]; pL[g] = PL[(size_t)g * ncol + q]; qF[g] = QF[(size_t)g * ncol + q]; qL[g] = QL[(size_t)g * ncol + q];
      yF[g] = W[(size_t)ncol * g * n3l + q]; yL[g] = W[(size_t)ncol * (g * n3l + n3l - 1) + q];
    }
    if (periodic) {
      dz_interface_solve(G, pF, pL, qF, qL, yF, yL, u);
      for (int g = 0; g < G; ++g) { XP[(size_t)g * ncol + q] = u[2 * ((g + G - 1) % G) + 1]; XN[(size_t)g * ncol + q] = u[2 * ((g + 1) % G)]; }
    } else {                                              // same branch as dz_interface_kernel; cross-checked against the dense solve
      double xp[FB_DZ_MAXG], xn[FB_DZ_MAXG];
      dz_interface_solve_walls(G, pF, pL, qF, qL, yF, yL, xp, xn);
      dz_interface_solve(G, pF, pL, qF, qL, yF, yL, u);
      for (int g = 0; g < G; ++g) {
        XP[(size_t)g * ncol + q] = xp[g]; XN[(size_t)g * ncol + q] = xn[g];
        const double dp = (g > 0) ? u[2 * (g - 1) + 1] : 0.0, dn = (g < G - 1) ? u[2 * (g + 1)] : 0.0;
        const double sc = 1e-9 * (fabs(dp) + fabs(dn) + 1e-300);
        if (fabs(xp[g] - dp) > sc + 1e-9 * fabs(dp) || fabs(xn[g] - dn) > sc + 1e-9 * fabs(dn)) return 9;
      }
    }
  }
  // pass 2: x = y + T_g^{-1}(-ca x_prev e_first - cc x_next e_last)   (the CORR variant of the kernel)
  for (int g = 0; g < G; ++g) {
    std::fill(scratch.begin(), scratch.end(), 0.0);
    for (long q = 0; q < ncol; ++q) {
      scratch[q] = -ca[g] * XP[(size_t)g * ncol + q];
      scratch[(size_t)ncol * (n3l - 1) + q] += -cc[g] * XN[(size_t)g * ncol + q];
    }
    g_dz_a = a + g * n3l; g_dz_b = b + g * n3l; g_dz_c = c + g * n3l;
    if (!dz_local_solve(n3l, ncol, uni[g], (singular && g == G - 1) ? 1 : 0, lam, scratch.data())) return 3;
    double* y = W + (size_t)ncol * g * n3l;
    for (size_t i = 0; i < (size_t)ncol * n3l; ++i) y[i] += scratch[i];
  }
  return 0;
}

extern "C" int emul_dz_general(int nz, int G, long ncol, int periodic, int singular, const double* a, const double* b, const double* c,
                               const double* lam, double* W) {
  g_dz_general = 1;
  const int rc = emul_dz(nz, G, ncol, periodic, singular, a, b, c, lam, W);
  g_dz_general = 0;
  return rc;
}
