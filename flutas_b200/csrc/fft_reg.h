// fft_reg.h -- entry points of the register-resident transform kernels (defined in fft_reg_{x,y}_{fwd,bwd}.cu).
#pragma once
#include <cuda_runtime.h>

#include "geom.cuh"
#include "reg_fft.cuh"

#include <cstdlib>

namespace fb {
// y tiles at N >= 1024 (a line needs >= 32 threads): "wide" = 16 lanes (128-byte row pieces) in 512-thread blocks, one
// block per SM; otherwise 8 lanes in 256-thread blocks, two blocks per SM.  Measured on one B200 at 1024^3 with the pair-pass
// kernels (profiles/r02_v_ab.log): forward narrow 4.21 / wide 3.76 ms, backward narrow 4.18 / wide 5.13 ms (1024 x 1024 x 512
// DCT lines: 2.09 / 1.85 and 2.15 / 2.57 ms) -> the forward kernel runs wide tiles from N = 1024 up, the backward one only at
// N = 2048 (where narrow tiles would be 32-byte pieces) or when the slab solver asks for them because the spectral side
// is written straight into peer memory (NVLink stores: 128-byte pieces move ~1.5x faster than 64-byte ones).
// FLUTAS_B200_YWIDE=0/1 overrides all of it.
inline int& y_wide_request() { static int v = 0; return v; }
inline bool y_wide_enabled() {
  static int env = -2;
  if (env == -2) { const char* e = getenv("FLUTAS_B200_YWIDE"); env = e ? (e[0] == '0' ? 0 : 1) : -1; }
  return env >= 0 ? env == 1 : y_wide_request() == 1;
}
inline bool y_wide_forced_off() { const char* e = getenv("FLUTAS_B200_YWIDE"); return e && e[0] == '0'; }
cudaError_t reg_run_x_fwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st);
cudaError_t reg_run_x_bwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st);
cudaError_t reg_run_y_fwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st);
cudaError_t reg_run_y_bwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st);
}  // namespace fb
