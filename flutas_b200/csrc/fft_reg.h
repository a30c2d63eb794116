// fft_reg.h -- entry points of the register-resident transform kernels (defined in fft_reg_{x,y}_{fwd,bwd}.cu).
#pragma once
#include <cuda_runtime.h>

#include "geom.cuh"
#include "reg_fft.cuh"

namespace fb {
cudaError_t reg_run_x_fwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st);
cudaError_t reg_run_x_bwd(const RegPlan& P, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale, int nsm,
                          cudaStream_t st);
cudaError_t reg_run_y_fwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st);
cudaError_t reg_run_y_bwd(const RegPlan& P, double* W, int n1, long n3, const SpecGeom& sg, int nsm, cudaStream_t st);
}  // namespace fb
