#include "fft_reg.cuh"
#include "fft_reg.h"
namespace fb {
cudaError_t reg_run_y_fwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st) {
  // N = 2048: 8 lanes x 64 threads in one 512-thread block (64-byte pieces) beat 4 lanes (32-byte pieces) on one GPU too:
  // 3.06 / 3.34 ms against 3.99 / 4.34 ms (fwd / bwd) at 2048 x 2048 x 128 on B200
  const bool wide = y_wide_enabled() || (P.N >= 1024 && !y_wide_forced_off());   // forward: 16-lane tiles from N = 1024 up (fft_reg.h)
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
