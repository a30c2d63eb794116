#include "fft_reg.cuh"
#include "fft_reg.h"
namespace fb {
cudaError_t reg_run_y_bwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st) {
  const bool wide = y_wide_enabled();
  switch (P.N) {
    case 32: return reg_launch_y<32, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 64: return reg_launch_y<64, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 128: return reg_launch_y<128, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 256: return reg_launch_y<256, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 512: return reg_launch_y<512, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 1024: return reg_launch_y<1024, false>(P, W, n1, n3, sg, nsm, wide, st);
    case 2048: return reg_launch_y<2048, false>(P, W, n1, n3, sg, nsm, wide, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
