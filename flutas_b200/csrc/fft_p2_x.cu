#include "fft_p2.cuh"
#include "fft_p2.h"
namespace fb {
cudaError_t p2_run_x(bool fwd, const LinePlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale,
                     cudaStream_t st) {
  switch (P.N) {
    case 64: return p2_launch_x<64, 16>(fwd, P, src, gs, dst, gd, scale, st);
    case 128: return p2_launch_x<128, 16>(fwd, P, src, gs, dst, gd, scale, st);
    case 256: return p2_launch_x<256, 16>(fwd, P, src, gs, dst, gd, scale, st);
    case 512: return p2_launch_x<512, 16>(fwd, P, src, gs, dst, gd, scale, st);
    case 1024: return p2_launch_x<1024, 8>(fwd, P, src, gs, dst, gd, scale, st);
    case 2048: return p2_launch_x<2048, 8>(fwd, P, src, gs, dst, gd, scale, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace fb
