#include "fft_reg.cuh"
#include "fft_reg.h"
namespace fb {
cudaError_t reg_run_y_fwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st) {
  const bool wide = y_wide_enabled();
  switch (P.N) {
    case 32: return reg_launch_y<32, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 64: return reg_launch_y<64, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 128: return reg_launch_y<128, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 256: return reg_launch_y<256, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 512: return reg_launch_y<512, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 1024: return reg_launch_y<1024, true>(P, W, n1, n3, sg, nsm, wide, st);
    case 2048: return reg_launch_y<2048, true>(P, W, n1, n3, sg, nsm, wide, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
