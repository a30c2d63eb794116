// CPU emulation of the tile-FFT core (flutas_b200/csrc/tile_fft.cuh) for the "not gpu" tests.
// TEST INFRASTRUCTURE: runs the kernels' per-thread phase functions in a serial loop over
// (lane, worker) with the phase boundaries where the CUDA kernels have __syncthreads().  It lets
// the index logic (Makhoul permutation, rotation swizzle, digit reversal, split/merge, mode map) be
// checked against the oracle without a GPU.  Nothing in the product links this file.
#include <cstdlib>
#include <cstring>
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
  if (fwd) {
    for (int q = 0; q < P.npass; ++q)
      phase([&](int lane, int w) { fft_pass<TB, ROT, true>(tile.data(), P, q, lane, w, nworkers); });
    phase([&](int lane, int w) { split_fwd<TB, ROT>(tile.data(), P, lane, w, nworkers); });
  } else {
    phase([&](int lane, int w) { merge_bwd<TB, ROT>(tile.data(), P, lane, w, nworkers); });
    for (int q = P.npass - 1; q >= 0; --q)
      phase([&](int lane, int w) { fft_pass<TB, ROT, false>(tile.data(), P, q, lane, w, nworkers); });
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
