#include "fft_p2.cuh"
#include "fft_p2.h"
namespace fb {
cudaError_t p2_run_y(bool fwd, const LinePlan& P, double* W, int n1, long n3, const SpecGeom& sg, cudaStream_t st) {
  switch (P.N) {
    case 64: return p2_launch_y<64, 16>(fwd, P, W, n1, n3, sg, st);
    case 128: return p2_launch_y<128, 16>(fwd, P, W, n1, n3, sg, st);
    case 256: return p2_launch_y<256, 16>(fwd, P, W, n1, n3, sg, st);
    case 512: return p2_launch_y<512, 16>(fwd, P, W, n1, n3, sg, st);
    case 1024: return p2_launch_y<1024, 8>(fwd, P, W, n1, n3, sg, st);
    case 2048: return p2_launch_y<2048, 8>(fwd, P, W, n1, n3, sg, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
