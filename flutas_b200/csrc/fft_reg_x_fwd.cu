#include "fft_reg.cuh"
#include "fft_reg.h"
namespace fb {
cudaError_t reg_run_x_fwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st) {
  switch (P.N) {
    case 32: return reg_launch_x<32, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 64: return reg_launch_x<64, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 128: return reg_launch_x<128, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 256: return reg_launch_x<256, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 512: return reg_launch_x<512, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 1024: return reg_launch_x<1024, true>(P, src, gs, dst, gd, scale, nsm, st);
    case 2048: return reg_launch_x<2048, true>(P, src, gs, dst, gd, scale, nsm, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
