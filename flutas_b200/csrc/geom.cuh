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

// (32-bit division: the number of lines of one rank is far below 2^31 -- checked by the launchers; a 64-bit division is a
// ~100-instruction subroutine, executed per line and per thread)
__device__ __forceinline__ long line_offset(const LineGeom& g, long line) {
  const unsigned ul = (unsigned)line, uk = ul / (unsigned)g.n2;
  const long k = (long)uk, j = (long)(ul - uk * (unsigned)g.n2);
  return g.off0 + j * g.sj + k * g.sk;
}

// Spectral side of the y stage.  On one GPU this is the work array itself; on a z-slab decomposition
// it is the x-split "pencil" layout of the exchange: element (i, row, k) of the slab lives in chunk
// q = i / n1l at ptr[q][koff + (i - q*n1l) + n1l*(row + N*k)].  ptr[q] is a send/receive buffer (NCCL
// path) or rank q's pencil buffer mapped through CUDA IPC (direct NVLink stores/loads).
// Reference counterpart: the pack/unpack loops of transpose_x_to_z / transpose_z_to_x
// (src/2decomp/transpose_x_to_z.f90:45-142, transpose_z_to_x.f90:15-163), fused here into the kernels.
enum { FB_MAX_RANKS = 8 };
struct SpecGeom {
  double* ptr[FB_MAX_RANKS];
  int n1l;       // x rows per chunk (n1 on one GPU)
  long koff;     // offset of this rank's k-range inside the destination chunk
};
__device__ __forceinline__ double* spec_base(const SpecGeom& g, int i, int N, long k) {
  const int q = i / g.n1l;
  return g.ptr[q] + g.koff + (i - q * g.n1l) + (long)g.n1l * N * k;
}

// Output side of the z stage: column `col`, level k goes to chunk q = k / n3l at
// ptr[q][koff + col + ncol*(k - q*n3l)] (one GPU: ptr[0] = the work array, n3l = nz, koff = 0).
struct ColGeom {
  double* ptr[FB_MAX_RANKS];
  int n3l;
  long koff;
};

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
