// fft_p2.cuh -- transform kernels specialised at compile time for power-of-two lengths (64..2048):
// radix schedule, strides and twiddle steps are constants, so the per-butterfly integer work of the
// generic kernels (half of their instructions in the first ncu profile) folds into immediates.
//
//   xfft_p2_kernel : x lines (contiguous): global <-> smem tile (transposing, rotation swizzle), passes
//                    and split/merge in shared memory.
//   yfft_p2_kernel : y lines (stride n1): the first pass reads its butterfly inputs straight from global
//                    memory (lane-contiguous 128/64-byte rows) and the split step writes the spectrum
//                    straight back (and the reverse for the inverse), so each element makes two fewer
//                    trips through shared memory and two __syncthreads disappear.
//
// Same arithmetic as the generic kernels (same pass_core/split_core/merge_core), hence identical results.
#pragma once
#include <cuda_runtime.h>

#include "geom.cuh"
#include "tile_fft.cuh"

namespace fb {

template <int M, int Q, bool FWD, int TB, bool ROT>
__device__ __forceinline__ void p2_tile_pass(double* tile, const cpx* wM, int lane, int worker, int nworkers) {
  if constexpr (Q < p2_npass(M)) {
    const TileAcc<TB, ROT> acc{tile, M, lane};
    p2_pass<M, Q, FWD>(wM, worker, nworkers, acc, acc);
    __syncthreads();
  }
}

template <int N, int TB, bool FWD>
__global__ void __launch_bounds__(256, (N * TB * 8 <= 72 * 1024) ? 3 : 1)
xfft_p2_kernel(LinePlan P, const double* __restrict__ src, LineGeom gs, double* __restrict__ dst, LineGeom gd,
               double scale) {
  constexpr int M = N / 2;
  extern __shared__ double tile[];
  const int kind = P.kind;
  cpx* s_w = reinterpret_cast<cpx*>(tile + (size_t)N * TB);
  long* s_off = reinterpret_cast<long*>(s_w + M);
  const int tid = threadIdx.x;
  constexpr int nthr = 256;
  const long line0 = (long)blockIdx.x * TB;
  const int nlive = (int)min((long)TB, gs.nlines - line0);

  if (tid < TB) s_off[tid] = line_offset(gs, min(line0 + tid, gs.nlines - 1));
  stage_twiddles(s_w, P.wM, M, tid, nthr);
  __syncthreads();

#pragma unroll 1
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

  const int lane = tid & (TB - 1), worker = tid / TB;
  constexpr int NW = nthr / TB;
  const TileAcc<TB, true> acc{tile, M, lane};
  if (FWD) {
    p2_tile_pass<M, 0, true, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 1, true, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 2, true, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 3, true, TB, true>(tile, s_w, lane, worker, NW);
    split_core(M, kind, P.wN, P.wQ, P.pos, worker, NW, acc, acc);
    __syncthreads();
  } else {
    merge_core(M, kind, P.wN, P.wQ, P.pos, worker, NW, acc, acc);
    __syncthreads();
    p2_tile_pass<M, 3, false, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 2, false, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 1, false, TB, true>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 0, false, TB, true>(tile, s_w, lane, worker, NW);
  }

  if (tid < TB) s_off[tid] = line_offset(gd, min(line0 + tid, gd.nlines - 1));
  __syncthreads();
#pragma unroll 1
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

template <int N, int TB, bool FWD>
__global__ void __launch_bounds__(256, (N * TB * 8 <= 72 * 1024) ? 3 : 1)
yfft_p2_kernel(LinePlan P, double* __restrict__ W, int n1, int ntile_i, SpecGeom sg) {
  constexpr int M = N / 2, NP = p2_npass(M);
  extern __shared__ double tile[];
  const int kind = P.kind;
  cpx* s_w = reinterpret_cast<cpx*>(tile + (size_t)N * TB);
  const int tid = threadIdx.x;
  constexpr int nthr = 256, NW = nthr / TB;
  const int ti = blockIdx.x % ntile_i;
  const long k = blockIdx.x / ntile_i;
  const int i0 = ti * TB;
  const int lane = tid & (TB - 1), worker = tid / TB;
  const bool live = (i0 + lane) < n1;
  const int il = i0 + (live ? lane : 0);
  double* base = W + (long)n1 * N * k + il;               // physical side
  double* sbase = spec_base(sg, il, N, k);                 // spectral side (pencil chunk of the exchange)

  stage_twiddles(s_w, P.wM, M, tid, nthr);
  __syncthreads();
  const TileAcc<TB, false> acc{tile, M, lane};
  if (FWD) {
    const LineAcc gin{base, (long)n1, kind, N, live, 1.0};
    p2_pass<M, 0, true>(s_w, worker, NW, gin, acc);           // global -> butterfly -> smem
    __syncthreads();
    p2_tile_pass<M, 1, true, TB, false>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 2, true, TB, false>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 3, true, TB, false>(tile, s_w, lane, worker, NW);
    const SpecAcc gout{sbase, (long)sg.n1l, M, live};
    split_core(M, kind, P.wN, P.wQ, P.pos, worker, NW, acc, gout);   // smem -> split -> global
  } else {
    const SpecAcc gin{sbase, (long)sg.n1l, M, live};
    merge_core(M, kind, P.wN, P.wQ, P.pos, worker, NW, gin, acc);    // global -> merge -> smem
    __syncthreads();
    p2_tile_pass<M, 3, false, TB, false>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 2, false, TB, false>(tile, s_w, lane, worker, NW);
    p2_tile_pass<M, 1, false, TB, false>(tile, s_w, lane, worker, NW);
    const LineAcc gout{base, (long)n1, kind, N, live, 1.0};
    p2_pass<M, 0, false>(s_w, worker, NW, acc, gout);          // smem -> butterfly -> global
  }
  (void)NP;
}

template <int N, int TB>
inline cudaError_t p2_launch_x(bool fwd, const LinePlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd,
                               double scale, cudaStream_t st) {
  const size_t smem = fft_smem_bytes<TB>(N);
  const long nblk = (gs.nlines + TB - 1) / TB;
  cudaError_t e;
  if (fwd) {
    e = cudaFuncSetAttribute(xfft_p2_kernel<N, TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    xfft_p2_kernel<N, TB, true><<<(unsigned)nblk, 256, smem, st>>>(P, src, gs, dst, gd, scale);
  } else {
    e = cudaFuncSetAttribute(xfft_p2_kernel<N, TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    xfft_p2_kernel<N, TB, false><<<(unsigned)nblk, 256, smem, st>>>(P, src, gs, dst, gd, scale);
  }
  return cudaGetLastError();
}

template <int N, int TB>
inline cudaError_t p2_launch_y(bool fwd, const LinePlan& P, double* W, int n1, long n3, const SpecGeom& sg,
                               cudaStream_t st) {
  const size_t smem = fft_smem_bytes<TB>(N);
  const int nti = (n1 + TB - 1) / TB;
  cudaError_t e;
  if (fwd) {
    e = cudaFuncSetAttribute(yfft_p2_kernel<N, TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    yfft_p2_kernel<N, TB, true><<<(unsigned)(nti * n3), 256, smem, st>>>(P, W, n1, nti, sg);
  } else {
    e = cudaFuncSetAttribute(yfft_p2_kernel<N, TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    yfft_p2_kernel<N, TB, false><<<(unsigned)(nti * n3), 256, smem, st>>>(P, W, n1, nti, sg);
  }
  return cudaGetLastError();
}

}  // namespace fb
