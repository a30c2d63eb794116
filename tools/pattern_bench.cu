// pattern_bench.cu -- access-pattern ceilings for the strided stages (measurement tool, not product code).
// Reads and writes an FP64 field the way the z solve / y transforms do -- tiles of TI consecutive elements x R rows,
// rows `rstride` elements apart -- with no arithmetic, so the result is the HBM rate the PATTERN allows.
//   z pattern: tile = TI columns x nz levels, row stride n1*n2            (thomas_reg_kernel)
//   y pattern: tile = TI lanes x n2 rows of one k-plane, row stride n1     (yfft_reg_kernel)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pattern_bench pattern_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

template <int LM> __device__ __forceinline__ double ldm(const double* p) {
  double x;
  if (LM == 0) x = *p;
  else if (LM == 1) x = __ldcs(p);
  else if (LM == 2) asm volatile("ld.global.L2::128B.f64 %0, [%1];" : "=d"(x) : "l"(p));
  else if (LM == 3) asm volatile("ld.global.L2::256B.f64 %0, [%1];" : "=d"(x) : "l"(p));
  else if (LM == 4) asm volatile("ld.global.cs.L2::256B.f64 %0, [%1];" : "=d"(x) : "l"(p));
  else asm volatile("ld.global.L1::no_allocate.L2::256B.f64 %0, [%1];" : "=d"(x) : "l"(p));
  return x;
}
template <int SM> __device__ __forceinline__ void stm(double* p, double v) {
  if (SM == 0) *p = v; else if (SM == 1) __stcs(p, v); else __stcg(p, v);
}

// tile t -> base offset: z: t*TI ; y: (t % nti)*TI + (t / nti) * plane
template <int TI, int L, int LM, int SM, int CL>
__global__ void __launch_bounds__(512) tile_copy(long rstride, int rows, long ntiles, int nti, long plane,
                                                 const double* __restrict__ in, double* __restrict__ out) {
  const int lane = threadIdx.x % TI, s = threadIdx.x / TI;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long off = (tile % nti) * TI + (tile / nti) * plane + lane + (long)(s * L) * rstride;
    const double* src = in + off;
    double* dst = out + off;
    double v[L];
#pragma unroll
    for (int l = 0; l < L; ++l) v[l] = ldm<LM>(src + (long)l * rstride);
#pragma unroll
    for (int l = 0; l < L; ++l) stm<SM>(dst + (long)l * rstride, v[l] + 1.0);
    if (CL > 1) cg::this_cluster().sync();
  }
}

template <int TI, int L, int LM, int SM, int CL>
float run(long rstride, int rows, long ntiles, int nti, long plane, const double* in, double* out, int bps, int reps) {
  const int threads = TI * (rows / L);
  if (threads > 512 || threads < 32) return -1.f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148 * bps;
  grid -= grid % CL;
  // every block must run the same number of iterations when it takes part in cluster barriers
  const long per = (ntiles / grid) * grid;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  auto kern = tile_copy<TI, L, LM, SM, CL>;
  for (int i = 0; i < 2; ++i) cudaLaunchKernelEx(&cfg, kern, rstride, rows, per, nti, plane, in, out);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, kern, rstride, rows, per, nti, plane, in, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess) return -2.f;
  return ms / reps * (float)ntiles / (float)per;
}

__global__ void plain_copy(long n, const double2* __restrict__ in, double2* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = in[i];
}

int main(int argc, char** argv) {
  const int n1 = argc > 1 ? atoi(argv[1]) : 512, n2 = argc > 2 ? atoi(argv[2]) : 512, n3 = argc > 3 ? atoi(argv[3]) : 512;
  const long ncol = (long)n1 * n2, npts = ncol * n3;
  double *a, *b;
  cudaMalloc(&a, npts * 8); cudaMalloc(&b, npts * 8);
  cudaMemset(a, 0, npts * 8); cudaMemset(b, 0, npts * 8);
  const double gb = 16.0 * npts / 1e9;
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    plain_copy<<<148 * 8, 256>>>(npts / 2, (double2*)a, (double2*)b);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) plain_copy<<<148 * 8, 256>>>(npts / 2, (double2*)a, (double2*)b);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid %d x %d x %d   plain copy: %.3f ms  %.0f GB/s\n", n1, n2, n3, ms / 5, gb / (ms / 5) * 1e3);
  }
#define Z(TI, L, LM, SM, CL, BPS) { float ms = run<TI, L, LM, SM, CL>(ncol, n3, ncol / TI, (int)(ncol / TI), 0, a, b, BPS, 5); \
    printf("z TI=%2d L=%2d ld=%d st=%d cluster=%d blocks/SM=%d threads=%4d : %7.3f ms  %5.0f GB/s\n", TI, L, LM, SM, CL, BPS, TI * (n3 / L), ms, gb / ms * 1e3); }
#define Y(TI, L, LM, SM, CL, BPS) { float ms = run<TI, L, LM, SM, CL>(n1, n2, (long)(n1 / TI) * n3, n1 / TI, ncol, a, b, BPS, 5); \
    printf("y TI=%2d L=%2d ld=%d st=%d cluster=%d blocks/SM=%d threads=%4d : %7.3f ms  %5.0f GB/s\n", TI, L, LM, SM, CL, BPS, TI * (n2 / L), ms, gb / ms * 1e3); }
  if (n3 <= 512) {
    Z(8, 16, 1, 1, 1, 2) Z(8, 16, 0, 0, 1, 2) Z(8, 16, 0, 1, 1, 2) Z(8, 16, 1, 0, 1, 2) Z(8, 16, 2, 0, 1, 2) Z(8, 16, 3, 0, 1, 2) Z(8, 16, 4, 1, 1, 2) Z(8, 16, 5, 0, 1, 2) Z(8, 16, 3, 2, 1, 2)
    Z(8, 16, 1, 1, 2, 2) Z(8, 16, 0, 0, 2, 2) Z(8, 16, 1, 1, 4, 2) Z(8, 16, 0, 0, 4, 2) Z(8, 16, 3, 0, 4, 2)
    Z(16, 16, 1, 1, 1, 1) Z(16, 16, 0, 0, 1, 1) Z(16, 16, 3, 0, 1, 1) Z(16, 16, 0, 0, 2, 1) Z(16, 16, 3, 0, 2, 1)
  } else {
    Z(8, 32, 1, 1, 1, 2) Z(8, 32, 0, 0, 1, 2) Z(8, 32, 2, 0, 1, 2) Z(8, 32, 3, 0, 1, 2) Z(8, 32, 4, 1, 1, 2) Z(8, 32, 5, 0, 1, 2) Z(8, 32, 3, 2, 1, 2)
    Z(8, 32, 1, 1, 2, 2) Z(8, 32, 0, 0, 2, 2) Z(8, 32, 1, 1, 4, 2) Z(8, 32, 0, 0, 4, 2) Z(8, 32, 3, 0, 4, 2) Z(8, 32, 3, 0, 8, 2)
    Z(16, 32, 1, 1, 1, 1) Z(16, 32, 0, 0, 1, 1) Z(16, 32, 3, 0, 1, 1) Z(16, 32, 0, 0, 2, 1) Z(16, 32, 3, 0, 2, 1)
  }
  if (n2 <= 512) {
    Y(16, 32, 1, 1, 1, 2) Y(16, 32, 0, 0, 1, 2) Y(16, 32, 3, 0, 1, 2) Y(16, 32, 0, 0, 2, 2) Y(16, 32, 3, 0, 2, 2) Y(8, 32, 1, 1, 1, 4) Y(8, 32, 3, 0, 1, 4) Y(32, 32, 1, 1, 1, 1) Y(32, 32, 0, 0, 1, 1)
  } else {
    Y(8, 32, 1, 1, 1, 2) Y(8, 32, 0, 0, 1, 2) Y(8, 32, 0, 1, 1, 2) Y(8, 32, 2, 0, 1, 2) Y(8, 32, 3, 0, 1, 2) Y(8, 32, 3, 1, 1, 2) Y(8, 32, 0, 0, 2, 2) Y(8, 32, 3, 0, 2, 2) Y(8, 32, 3, 0, 4, 2)
    Y(16, 64, 1, 1, 1, 2) Y(16, 64, 0, 0, 1, 2) Y(16, 64, 3, 0, 1, 2) Y(16, 32, 1, 1, 1, 1) Y(16, 32, 3, 0, 1, 1)
  }
  return 0;
}
