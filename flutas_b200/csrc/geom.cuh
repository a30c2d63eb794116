// geom.cuh -- line geometry and shared-memory helpers shared by the generic and the power-of-two kernels.
#pragma once
#include <cuda_runtime.h>

#include "tile_fft.cuh"

namespace fb {

// Where a set of x-lines lives: line l = j + n2*k  ->  base[off0 + j*sj + k*sk + i]
struct LineGeom {
  long off0, sj, sk;
  int n2;
  long nlines;
};

__device__ __forceinline__ long line_offset(const LineGeom& g, long line) {
  const long k = line / g.n2, j = line - k * g.n2;
  return g.off0 + j * g.sj + k * g.sk;
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout of the transform kernels: [tile: N*TB doubles][wM: M cpx][line offsets: TB longs]
// The pass twiddles wM are staged in shared memory (they are read 7x per radix-8 butterfly and global
// loads of them were the top stall in the first profile); wN/wQ/pos stay in global memory (L1/L2 hits).
template <int TB>
__host__ __device__ inline size_t fft_smem_bytes(int N) {
  return (size_t)N * TB * sizeof(double) + (size_t)(N / 2) * sizeof(cpx) + TB * sizeof(long);
}

__device__ __forceinline__ void stage_twiddles(cpx* s_w, const cpx* __restrict__ g_w, int M, int tid, int nthr) {
  const double2* g = reinterpret_cast<const double2*>(g_w);
  double2* d = reinterpret_cast<double2*>(s_w);
  for (int q = tid; q < M; q += nthr) d[q] = __ldg(g + q);
}

}  // namespace fb
