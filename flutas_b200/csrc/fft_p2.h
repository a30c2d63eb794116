// fft_p2.h -- entry points of the power-of-two transform kernels (defined in fft_p2_x.cu / fft_p2_y.cu).
#pragma once
#include <cuda_runtime.h>

#include "geom.cuh"

namespace fb {
// lengths served by the specialised kernels and the tile width each uses
inline int p2_tile_width(int N) {
  switch (N) {
    case 64: case 128: case 256: case 512: return 16;
    case 1024: case 2048: return 8;
    default: return 0;
  }
}
cudaError_t p2_run_x(bool fwd, const LinePlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale,
                     cudaStream_t st);
cudaError_t p2_run_y(bool fwd, const LinePlan& P, double* W, int n1, long n3, const SpecGeom& sg, cudaStream_t st);
}  // namespace fb
