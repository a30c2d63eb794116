// line_plan.h -- host-side construction of a LinePlan: radix schedule, twiddle tables, digit-reversal
// table and the row -> FFTW-mode map used to permute lambdaxy at plan time.
// Plain C++ (no CUDA) so tests/emulate can build it with g++.
//
// Reference counterparts: plan creation in fftini (src/fft.f90:64-157) and the BC -> transform table
// of find_fft (src/fft.f90:233-291); the eigenvalue ordering that has to match the spectral layout is
// built in eigenvalues() (src/initsolver.f90:122-186).
#pragma once
#include <cmath>
#include <vector>

#include "tile_fft.cuh"

namespace fb {

struct HostLinePlan {
  int N = 0, M = 0, kind = 0;
  std::vector<int> radix, sub;
  std::vector<cpx> wM, wN, wQ;
  std::vector<int> pos;        // pos[k]  : row of complex mode k
  std::vector<int> mode;       // mode[r] : index into the reference's lambda array for tile row r (0-based)
  bool ok = false;
};

inline cpx unit_root(long double num, long double den) {   // exp(-i pi num/den)
  const long double PI_L = 3.14159265358979323846264338327950288L;
  const long double ang = -PI_L * num / den;
  cpx r; r.x = (double)cosl(ang); r.y = (double)sinl(ang);
  return r;
}

// BC pair -> kind; returns -1 if the transform is not available on this path
inline int kind_from_bc(char b0, char b1) {
  if (b0 == 'P' && b1 == 'P') return KIND_PP;
  if (b0 == 'N' && b1 == 'N') return KIND_NN;
  if (b0 == 'D' && b1 == 'D') return KIND_DD;
  if (b0 == 'N' && b1 == 'D') return KIND_ND;             // REDFT11 both ways (src/fft.f90:256-259)
  if (b0 == 'D' && b1 == 'N') return KIND_DN;             // RODFT11 both ways (:260-263)
  return -1;
}

inline HostLinePlan make_line_plan(int N, int kind) {
  HostLinePlan hp;
  hp.N = N; hp.M = N / 2; hp.kind = kind;
  if (N < 2 || (N & 1)) return hp;                        // FluTAS requires even ng (sanity.f90:155)
  int rem = hp.M;
  while (rem % 8 == 0) { hp.radix.push_back(8); rem /= 8; }
  while (rem % 4 == 0) { hp.radix.push_back(4); rem /= 4; }
  while (rem % 2 == 0) { hp.radix.push_back(2); rem /= 2; }
  while (rem % 3 == 0) { hp.radix.push_back(3); rem /= 3; }
  while (rem % 5 == 0) { hp.radix.push_back(5); rem /= 5; }
  if (rem != 1 || (int)hp.radix.size() > FB_MAX_PASS) return hp;   // other prime factors: unsupported
  const int M = hp.M;
  int prod = 1;
  for (int r : hp.radix) { prod *= r; hp.sub.push_back(M / prod); }
  hp.wM.resize(M > 0 ? M : 1);
  for (int k = 0; k < M; ++k) hp.wM[k] = unit_root(2.0L * k, (long double)M);
  if (kind_is_iv(kind)) {                                 // post-twiddle e^{-i pi k/N}, pre-twiddle e^{-i pi (4m+1)/(4N)}
    hp.wN.resize(M);
    for (int k = 0; k < M; ++k) hp.wN[k] = unit_root((long double)k, (long double)N);
    hp.wQ.resize(M + 1);
    for (int k = 0; k <= M; ++k) hp.wQ[k] = unit_root(4.0L * k + 1.0L, 4.0L * N);
  } else {
    hp.wN.resize(M / 2 + 1);
    for (int k = 0; k <= M / 2; ++k) hp.wN[k] = unit_root(2.0L * k, (long double)N);
    hp.wQ.resize(M + 1);
    for (int k = 0; k <= M; ++k) hp.wQ[k] = unit_root((long double)k, 2.0L * N);
  }
  hp.pos.resize(M);
  for (int k = 0; k < M; ++k) {                           // k = t0 + r0 (t1 + r1 (t2 + ...)) -> sum t_q sub_q
    int kk = k, p = 0;
    for (size_t q = 0; q < hp.radix.size(); ++q) { p += (kk % hp.radix[q]) * hp.sub[q]; kk /= hp.radix[q]; }
    hp.pos[k] = p;
  }
  hp.mode.resize(N);
  for (int k = 0; k < M; ++k) {
    const int r = hp.pos[k];
    int q0 = k, q1 = (k == 0) ? M : N - k;                // part 0 / part 1 content (see split_fwd)
    if (kind == KIND_DD) { q0 = N - 1 - q0; q1 = N - 1 - q1; }
    if (kind_is_iv(kind)) { q0 = 2 * k; q1 = N - 1 - 2 * k; }   // rows (Y_{2k}, Y_{N-1-2k}), see iv_post
    hp.mode[r] = q0;
    hp.mode[M + r] = q1;
  }
  hp.ok = true;
  return hp;
}

}  // namespace fb

// ---------------------------------------------------------------------------------------------
// tables of the register-resident transforms (reg_fft.cuh)
#include "reg_fft.cuh"

namespace fb {

struct HostRegPlan {
  int N = 0, M = 0, kind = 0;
  std::vector<cpx> tw[RF_MAXPASS];
  std::vector<cpx> tw8[RF_MAXPASS];   // pass twiddles of the 8-values-per-thread schedule (M = 512 only)
  std::vector<cpx> wN, wQ;
  std::vector<int> mode;       // mode[r] : index into the reference's lambda array for spectral row r (0-based)
  bool ok = false;
};

template <class S>
inline void fill_sched_twiddles(std::vector<cpx>* tw) {
  for (int q = 1; q < S::NP; ++q) {
    const int r = S::radix(q), Ns = S::ns(q);
    tw[q].resize((size_t)(r - 1) * Ns);
    for (int t = 1; t < r; ++t)
      for (int k = 0; k < Ns; ++k) tw[q][(size_t)(t - 1) * Ns + k] = unit_root(2.0L * t * k, (long double)Ns * r);
  }
}
template <int M>
inline void fill_reg_twiddles(HostRegPlan& hp) {
  fill_sched_twiddles<RegSched<M>>(hp.tw);
  if (M == 512) fill_sched_twiddles<RegSched<512, 8>>(hp.tw8);
}

inline HostRegPlan make_reg_plan(int N, int kind) {
  HostRegPlan hp;
  hp.N = N; hp.M = N / 2; hp.kind = kind;
  if (!reg_fft_supported(N)) return hp;
  switch (hp.M) {
    case 16: fill_reg_twiddles<16>(hp); break;
    case 32: fill_reg_twiddles<32>(hp); break;
    case 64: fill_reg_twiddles<64>(hp); break;
    case 128: fill_reg_twiddles<128>(hp); break;
    case 256: fill_reg_twiddles<256>(hp); break;
    case 512: fill_reg_twiddles<512>(hp); break;
    case 1024: fill_reg_twiddles<1024>(hp); break;
    default: return hp;
  }
  const int M = hp.M;
  hp.wN.resize(M);
  hp.wQ.resize(M + 1);
  if (kind_is_iv(kind)) {
    for (int k = 0; k < M; ++k) hp.wN[k] = unit_root((long double)k, (long double)N);
    for (int k = 0; k <= M; ++k) hp.wQ[k] = unit_root(4.0L * k + 1.0L, 4.0L * N);
  } else {
    for (int k = 0; k < M; ++k) hp.wN[k] = unit_root(2.0L * k, (long double)N);
    for (int k = 0; k <= M; ++k) hp.wQ[k] = unit_root((long double)k, 2.0L * N);
  }
  hp.mode.resize(N);
  for (int r = 0; r < N; ++r) hp.mode[r] = reg_mode_index(N, kind, r);
  hp.ok = true;
  return hp;
}

}  // namespace fb
