#include "fft_reg.cuh"
#include "fft_reg.h"
namespace fb {
cudaError_t reg_run_x_bwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st) {
  switch (P.N) {
    case 32: return reg_launch_x<32, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 64: return reg_launch_x<64, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 128: return reg_launch_x<128, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 256: return reg_launch_x<256, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 512: return reg_launch_x<512, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 1024: return reg_launch_x<1024, false>(P, src, gs, dst, gd, scale, nsm, st);
    case 2048: return reg_launch_x<2048, false>(P, src, gs, dst, gd, scale, nsm, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
