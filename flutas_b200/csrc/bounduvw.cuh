// bounduvw.cuh -- ghost cells of the velocity field and the CFL reduction on the device.
//
//   set_bc_kernel      set_bc, src/bound.f90:227-646, any halo width: periodic wrap, cell-centred D/N, and the
//                      face-centred (wall-normal) D/N closures of a staggered component
//   outflow_kernel     outflow, src/bound.f90:649-773: wall-normal face velocity from zero divergence
//   chkdt_kernel       the field reduction of chkdt_sp / chkdt_tw, src/chkdt.f90:62-85 = :150-173 (max of the three
//                      convective inverse time scales); the scalar formulas that follow stay on the host
//
// One thread owns one line along the boundary-normal direction and walks the halo layers q = 0..nh-1 in the reference's
// order (a later q may overwrite an earlier one: face-centred D at the upper wall, :432-533), so the result is bit-identical
// with the Fortran loops whatever the thread schedule.  Arithmetic is evaluated as written, one rounding per operation.
#pragma once
#include <cuda_runtime.h>

namespace fb {

enum { FB_MAX_HALO = 8 };

// a field dimensioned (1-nh:n1+nh, 1-nh:n2+nh, 1-nh:n3+nh)
struct HaloField {
  int n[3];
  int nh;
  long s[3];                     // element strides of the three dimensions
};
__host__ __device__ inline HaloField halo_field(int n1, int n2, int n3, int nh) {
  HaloField g;
  g.n[0] = n1; g.n[1] = n2; g.n[2] = n3; g.nh = nh;
  g.s[0] = 1; g.s[1] = n1 + 2 * nh; g.s[2] = (long)(n1 + 2 * nh) * (n2 + 2 * nh);
  return g;
}

struct BcFactor { double f[FB_MAX_HALO]; };

enum BcMode { BC_WRAP = 0, BC_CENTERED = 1, BC_FACE_D = 2, BC_FACE_N = 3 };

// set_bc(nx,ny,nz,ctype,ibound,idir,centered,rvalue,qq_d,nh_p,dr,p); idir 0-based.  The two tangential loops run over the
// full extent including halos, like the reference's (`do k=1-nh_p,nz+nh_p`), so edges and corners come out identical.
__global__ void set_bc_kernel(double* __restrict__ p, HaloField g, int idir, int ibound, int mode, double sgn, BcFactor fac) {
  const int da = (idir == 0) ? 1 : 0, db = (idir == 2) ? 1 : 2;
  const int ea = g.n[da] + 2 * g.nh, eb = g.n[db] + 2 * g.nh;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)ea * eb) return;
  const int a = (int)(idx % ea), b = (int)(idx / ea);
  double* line = p + a * g.s[da] + b * g.s[db];
  const long sd = g.s[idir];
  const int n = g.n[idir], nh = g.nh;
  auto at = [&](int i) -> double& { return line[(long)(i + nh - 1) * sd]; };     // Fortran index i of dimension idir
  if (mode == BC_WRAP) {                                  // :268-318
    for (int q = 0; q < nh; ++q) { at(0 - q) = at(n - q); at(n + 1 + q) = at(1 + q); }
    return;
  }
  for (int q = 0; q < nh; ++q) {
    const double f = fac.f[q];
    if (mode == BC_CENTERED) {                            // :320-431  factor_value+sgn*p
      if (ibound == 0) at(0 - q) = __dadd_rn(f, __dmul_rn(sgn, at(1 + q)));
      else at(n + 1 + q) = __dadd_rn(f, __dmul_rn(sgn, at(n - q)));
    } else if (mode == BC_FACE_D) {                       // :432-533
      if (ibound == 0) at(0 - q) = f;
      else { at(n + q) = f; at(n + 1 + q) = at(n - 1 - q); }
    } else {                                              // BC_FACE_N, :534-646
      if (ibound == 0) at(0 - q) = __dadd_rn(__dmul_rn(1.0, f), at(1 + q));
      else {
        at(n + q) = __dadd_rn(__dmul_rn(1.0, f), at(n - 1 - q));
        at(n + 1 + q) = __dadd_rn(__dmul_rn(2.0, f), at(n - 1 - q));
      }
    }
  }
}

// outflow(nx,ny,nz,idir,nh_d,nh_u,dx,dy,dz,dzf,u,v,w), src/bound.f90:649-773; dir = +-(1,2,3) as in the reference.
// dzf points at index 1-nh_d.  One thread per point of the boundary face (interior range of the two tangential directions).
__global__ void outflow_kernel(HaloField g, int dir, int nh_d, double dx, double dy, const double* __restrict__ dzf,
                               double* __restrict__ u, double* __restrict__ v, double* __restrict__ w) {
  const int ad = (dir < 0 ? -dir : dir) - 1;              // 0-based normal direction
  const int da = (ad == 0) ? 1 : 0, db = (ad == 2) ? 1 : 2;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)g.n[da] * g.n[db]) return;
  const int a = (int)(idx % g.n[da]) + 1, b = (int)(idx / g.n[da]) + 1;       // Fortran indices 1..n
  const int nh = g.nh, qmin = nh - 1;
  const double dxi = __ddiv_rn(1.0, dx), dyi = __ddiv_rn(1.0, dy);            // dx**(-1), dy**(-1)
  auto off = [&](int i, int j, int k) { return (long)(i + nh - 1) + (long)(j + nh - 1) * g.s[1] + (long)(k + nh - 1) * g.s[2]; };
  auto zf = [&](int k) { return dzf[k + nh_d - 1]; };
  if (ad == 0) {
    const int j = a, k = b;
    const double dzk = __ddiv_rn(1.0, zf(k));              // dzfi(k) = dzf(k)**(-1)
    if (dir > 0) {
      const int i = g.n[0];
      for (int q = 0; q <= qmin + 1; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(v[off(i + q, j, k)], v[off(i + q, j - 1, k)]), dyi),
                                   __dmul_rn(__dsub_rn(w[off(i + q, j, k)], w[off(i + q, j, k - 1)]), dzk));
        u[off(i + q, j, k)] = __dsub_rn(u[off(i - 1 - q, j, k)], __dmul_rn(dx, t));
      }
    } else {
      const int i = 0;
      for (int q = 0; q <= qmin; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(v[off(i + 1 + q, j, k)], v[off(i + 1 + q, j - 1, k)]), dyi),
                                   __dmul_rn(__dsub_rn(w[off(i + 1 + q, j, k)], w[off(i + 1 + q, j, k - 1)]), dzk));
        u[off(i - q, j, k)] = __dadd_rn(u[off(i + 1 + q, j, k)], __dmul_rn(dx, t));
      }
    }
  } else if (ad == 1) {
    const int i = a, k = b;
    const double dzk = __ddiv_rn(1.0, zf(k));
    if (dir > 0) {
      const int j = g.n[1];
      for (int q = 0; q <= qmin + 1; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(u[off(i, j + q, k)], u[off(i - 1, j + q, k)]), dxi),
                                   __dmul_rn(__dsub_rn(w[off(i, j + q, k)], w[off(i, j + q, k - 1)]), dzk));
        v[off(i, j + q, k)] = __dsub_rn(v[off(i, j - 1 - q, k)], __dmul_rn(dy, t));
      }
    } else {
      const int j = 0;
      for (int q = 0; q <= qmin; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(u[off(i, j + 1 + q, k)], u[off(i - 1, j + 1 + q, k)]), dxi),
                                   __dmul_rn(__dsub_rn(w[off(i, j + 1 + q, k)], w[off(i, j + 1 + q, k - 1)]), dzk));
        v[off(i, j - q, k)] = __dadd_rn(v[off(i, j + 1 + q, k)], __dmul_rn(dy, t));
      }
    }
  } else {
    const int i = a, j = b;
    if (dir > 0) {
      const int k = g.n[2];
      for (int q = 0; q <= qmin + 1; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(u[off(i, j, k + q)], u[off(i - 1, j, k + q)]), dxi),
                                   __dmul_rn(__dsub_rn(v[off(i, j, k + q)], v[off(i, j - 1, k + q)]), dyi));
        w[off(i, j, k + q)] = __dsub_rn(w[off(i, j, k - 1 - q)], __dmul_rn(zf(k + q), t));
      }
    } else {
      const int k = 0;
      for (int q = 0; q <= qmin; ++q) {
        const double t = __dadd_rn(__dmul_rn(__dsub_rn(u[off(i, j, k + 1 + q)], u[off(i - 1, j, k + 1 + q)]), dxi),
                                   __dmul_rn(__dsub_rn(v[off(i, j, k + 1 + q)], v[off(i, j - 1, k + 1 + q)]), dyi));
        w[off(i, j, k - q)] = __dadd_rn(w[off(i, j, k + 1 + q)], __dmul_rn(zf(k - q), t));
      }
    }
  }
}

// chkdt (src/chkdt.f90:150-173): per-block maximum of max(dtix, dtiy, dtiz); dzci, dzfi point at index 1-nh_d
__global__ void __launch_bounds__(256) chkdt_kernel(HaloField g, int nh_d, double dxi, double dyi,
                                                    const double* __restrict__ dzci, const double* __restrict__ dzfi,
                                                    const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, double* __restrict__ part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  const int nh = g.nh;
  double dti = 0.0;
  if (i <= g.n[0] && j <= g.n[1]) {
    auto off = [&](int ii, int jj, int kk) { return (long)(ii + nh - 1) + (long)(jj + nh - 1) * g.s[1] + (long)(kk + nh - 1) * g.s[2]; };
    auto U = [&](int a, int b, int c) { return u[off(i + a, j + b, k + c)]; };
    auto V = [&](int a, int b, int c) { return v[off(i + a, j + b, k + c)]; };
    auto W = [&](int a, int b, int c) { return w[off(i + a, j + b, k + c)]; };
    auto sum4 = [](double a, double b, double c, double d) { return __dadd_rn(__dadd_rn(__dadd_rn(a, b), c), d); };
    auto comb = [](double a, double ca, double b, double cb, double c, double cc) {
      return __dadd_rn(__dadd_rn(__dmul_rn(a, ca), __dmul_rn(b, cb)), __dmul_rn(c, cc));
    };
    const double zf = dzfi[k + nh_d - 1], zc = dzci[k + nh_d - 1];
    const double ux = fabs(U(0, 0, 0));
    const double vx = __dmul_rn(0.25, fabs(sum4(V(0, 0, 0), V(0, -1, 0), V(1, 0, 0), V(1, -1, 0))));
    const double wx = __dmul_rn(0.25, fabs(sum4(W(0, 0, 0), W(0, 0, -1), W(1, 0, 0), W(1, 0, -1))));
    const double dtix = comb(ux, dxi, vx, dyi, wx, zf);
    const double uy = __dmul_rn(0.25, fabs(sum4(U(0, 0, 0), U(0, 1, 0), U(-1, 1, 0), U(-1, 0, 0))));
    const double vy = fabs(V(0, 0, 0));
    const double wy = __dmul_rn(0.25, fabs(sum4(W(0, 0, 0), W(0, 1, 0), W(0, 1, -1), W(0, 0, -1))));
    const double dtiy = comb(uy, dxi, vy, dyi, wy, zf);
    const double uz = __dmul_rn(0.25, fabs(sum4(U(0, 0, 0), U(-1, 0, 0), U(-1, 0, 1), U(0, 0, 1))));
    const double vz = __dmul_rn(0.25, fabs(sum4(V(0, 0, 0), V(0, -1, 0), V(0, -1, 1), V(0, 0, 1))));
    const double wz = fabs(W(0, 0, 0));
    const double dtiz = comb(uz, dxi, vz, dyi, wz, zc);
    dti = fmax(fmax(dtix, dtiy), dtiz);
  }
  __shared__ double red[256];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  red[tid] = dti;
  __syncthreads();
  for (int sft = nt / 2; sft > 0; sft >>= 1) {
    if (tid < sft) red[tid] = fmax(red[tid], red[tid + sft]);
    __syncthreads();
  }
  if (tid == 0) part[((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = red[0];
}

__global__ void max_final_kernel(long nparts, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double red[256];
  double m = 0.0;
  for (long q = threadIdx.x; q < nparts; q += blockDim.x) m = fmax(m, part[q]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int sft = blockDim.x / 2; sft > 0; sft >>= 1) {
    if ((int)threadIdx.x < sft) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + sft]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

}  // namespace fb
