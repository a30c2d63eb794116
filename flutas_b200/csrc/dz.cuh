// dz.cuh -- distributed z solve: the z-slab decomposition WITHOUT the two all-to-all transposes.
//
// The reference's slab path (and round 1 of this library) transposes the whole field to x-split z-pencils, solves, and
// transposes back: 2 x 7/8 of the field crosses NVLink per solve (1.9 GB per GPU at 1024^3 on 8 GPUs) and those two
// exchanges are 2/3 of the 8-GPU solve time.  A tridiagonal system split over G ranks does not need that:
//
//   rank g owns levels [g n3l, (g+1) n3l) of EVERY column; T_g = its diagonal block of the z operator (couplings cut)
//   pass 1  y = T_g^{-1} b                       local, the ordinary z kernel on (ncol x n3l)            16 B/pt of HBM
//   exchange the two boundary planes y_first, y_last to the column's owner                                16 B/column
//   interface system per column (2G unknowns: first and last unknown of every rank)
//           first_g + pF_g last_{g-1} + qF_g first_{g+1} = yF_g ,   last_g + pL_g last_{g-1} + qL_g first_{g+1} = yL_g
//           with p = T_g^{-1}(a_first e_first), q = T_g^{-1}(c_last e_last) at the first / last level: right-hand-side
//           independent, computed once per plan with two local solves
//   send x_prev = last_{g-1}, x_next = first_{g+1} back                                                   16 B/column
//   pass 2  x = y + T_g^{-1}( -a_first x_prev e_first - c_last x_next e_last )     local (CORR kernel)    16 B/pt
//
// NVLink traffic per GPU and solve: 32 B per column instead of 2 x 8 B x 7/8 per POINT (1024^3, 8 GPUs: 29 MB
// instead of 1.9 GB); the price is one more HBM pass of the z stage (32 instead of 16 B/pt).  Periodic z couples rank 0
// to rank G-1 through the same formulas; the pinned singular column lives in the last rank's local block.
// The ill-conditioned columns (thomas_ref.cuh) are gathered whole on their owner, solved in the reference's order and
// written back over the result.
//
// Reference counterpart: transpose_xc_to_z / gaussel / transpose_z_to_xc (src/solver_gpu.f90:150-182).
// Host-compilable core (tests/emulate): dz_interface_solve.
#pragma once
#include <math.h>

#include "tile_fft.cuh"

namespace fb {

enum { FB_DZ_MAXG = 8 };

// Interface system of one column: unknowns u[2g] = first, u[2g+1] = last unknown of rank g (indices modulo G; the
// couplings that do not exist in a non-periodic system are zero in p / q already).  Dense elimination with partial
// pivoting on at most 16 x 16.
FB_HD void dz_interface_solve(int G, const double* pF, const double* pL, const double* qF, const double* qL, const double* yF,
                              const double* yL, double* u) {
  const int n = 2 * G;
  double M[2 * FB_DZ_MAXG][2 * FB_DZ_MAXG + 1];
  for (int r = 0; r < n; ++r) for (int c = 0; c <= n; ++c) M[r][c] = 0.0;
  for (int g = 0; g < G; ++g) {
    const int gl = (g + G - 1) % G, gr = (g + 1) % G;
    M[2 * g][2 * g] += 1.0;         M[2 * g][2 * gl + 1] += pF[g];     M[2 * g][2 * gr] += qF[g];     M[2 * g][n] = yF[g];
    M[2 * g + 1][2 * g + 1] += 1.0; M[2 * g + 1][2 * gl + 1] += pL[g]; M[2 * g + 1][2 * gr] += qL[g]; M[2 * g + 1][n] = yL[g];
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
    if (piv != c) for (int k = 0; k <= n; ++k) { const double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
    const double inv = 1.0 / M[c][c];
    for (int r = c + 1; r < n; ++r) {
      const double f = M[r][c] * inv;
      if (f != 0.0) for (int k = c; k <= n; ++k) M[r][k] -= f * M[c][k];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double acc = M[r][n];
    for (int k = r + 1; k < n; ++k) acc -= M[r][k] * u[k];
    u[r] = acc / M[r][r];
  }
}

// Non-periodic z: the same system, solved in O(G).  With w_j = (last_j, first_{j+1}), j = 0..G-2 (the two unknowns
// either side of the cut between rank j and rank j+1) the equations become block tridiagonal with 2 x 2 blocks,
//     [ 1        qL_j ] w_j  +  [ pL_j 0 ] w_{j-1}  +  [ 0 0        ] w_{j+1}  =  ( yL_j     )
//     [ pF_{j+1} 1    ]         [ 0    0 ]             [ 0 qF_{j+1} ]             ( yF_{j+1} )
// (pF_0 = pL_0 = 0: rank 0 has no neighbour below; qF_{G-1} = qL_{G-1} = 0).  Block Thomas: the only fill-in is one entry per
// block, the determinants 1 - qL pF stay positive because |p|, |q| < 1 for the (negative definite) z operator.
// Output: xprev[g] = last_{g-1}, xnext[g] = first_{g+1} (0 where the neighbour does not exist).
FB_HD void dz_interface_solve_walls(int G, const double* pF, const double* pL, const double* qF, const double* qL, const double* yF,
                                    const double* yL, double* xprev, double* xnext) {
  double b01[FB_DZ_MAXG], b10[FB_DZ_MAXG], r0[FB_DZ_MAXG], r1[FB_DZ_MAXG], idet[FB_DZ_MAXG];
  const int m = G - 1;                                     // number of cuts
  // forward elimination: B'_j = B_j - A_j B'^{-1}_{j-1} C_{j-1} touches entry (0,1) only; r'_j = r_j - A_j B'^{-1}_{j-1} r'_{j-1} entry 0 only
  for (int j = 0; j < m; ++j) {
    double e01 = qL[j], f0 = yL[j];
    const double e10 = pF[j + 1], f1 = yF[j + 1];
    if (j > 0) {
      // X = B'^{-1}_{j-1} = idet [[1, -b01],[-b10, 1]];  X[0][1] = -b01 idet;  (X r')[0] = idet (r0 - b01 r1)
      const double x01 = -b01[j - 1] * idet[j - 1];
      const double xr0 = idet[j - 1] * (r0[j - 1] - b01[j - 1] * r1[j - 1]);
      e01 -= pL[j] * x01 * qF[j];
      f0 -= pL[j] * xr0;
    }
    b01[j] = e01; b10[j] = e10; r0[j] = f0; r1[j] = f1;
    idet[j] = 1.0 / (1.0 - e01 * e10);
  }
  // back substitution: w_j = B'^{-1}_j (r'_j - C_j w_{j+1}),  C_j w_{j+1} = (0, qF_{j+1} first_{j+2})
  double wl = 0.0, wf = 0.0;
  for (int g = 0; g < G; ++g) { xprev[g] = 0.0; xnext[g] = 0.0; }
  for (int j = m - 1; j >= 0; --j) {
    const double s0 = r0[j], s1 = r1[j] - ((j + 1 < m) ? qF[j + 1] * wf : 0.0);
    const double nl = idet[j] * (s0 - b01[j] * s1);
    const double nf = idet[j] * (s1 - b10[j] * s0);
    wl = nl; wf = nf;
    xprev[j + 1] = wl;                                     // last_j is what rank j+1 calls x_prev
    xnext[j] = wf;                                         // first_{j+1} is what rank j calls x_next
  }
}

}  // namespace fb

#if defined(__CUDACC__)
namespace fb {

struct DzPeers { double* p[FB_DZ_MAXG]; };

// two planes of this rank's slab (levels first / last, ncol doubles each) -> the column owners: owner q = col / ncol_own
// receives them at dstF[q] + rank*ncol_own + (col - q ncol_own), likewise dstL (direct NVLink stores, 16 B per column)
__global__ void dz_send_planes_kernel(long ncol, long ncol_own, int rank, const double* __restrict__ first,
                                      const double* __restrict__ last, DzPeers dstF, DzPeers dstL) {
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int q = (int)(col / ncol_own);
  const long o = (long)rank * ncol_own + (col - (long)q * ncol_own);
  dstF.p[q][o] = first[col];
  dstL.p[q][o] = last[col];
}

// one thread per owned column: 2G x 2G interface system; the neighbours' boundary unknowns go straight to every rank
// (xprev[g][col], xnext[g][col] with col the global column index)
__global__ void dz_interface_kernel(int G, int periodic, long ncol_own, long col0, const double* __restrict__ PF, const double* __restrict__ PL,
                                    const double* __restrict__ QF, const double* __restrict__ QL, const double* __restrict__ YF,
                                    const double* __restrict__ YL, DzPeers xprev, DzPeers xnext) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol_own) return;
  double pF[FB_DZ_MAXG], pL[FB_DZ_MAXG], qF[FB_DZ_MAXG], qL[FB_DZ_MAXG], yF[FB_DZ_MAXG], yL[FB_DZ_MAXG], u[2 * FB_DZ_MAXG];
  for (int g = 0; g < G; ++g) {
    const long o = (long)g * ncol_own + c;
    pF[g] = PF[o]; pL[g] = PL[o]; qF[g] = QF[o]; qL[g] = QL[o]; yF[g] = YF[o]; yL[g] = YL[o];
  }
  if (periodic) {                                          // cyclic coupling rank 0 <-> rank G-1: dense elimination
    dz_interface_solve(G, pF, pL, qF, qL, yF, yL, u);
    for (int g = 0; g < G; ++g) {
      xprev.p[g][col0 + c] = u[2 * ((g + G - 1) % G) + 1];
      xnext.p[g][col0 + c] = u[2 * ((g + 1) % G)];
    }
  } else {
    double xp[FB_DZ_MAXG], xn[FB_DZ_MAXG];
    dz_interface_solve_walls(G, pF, pL, qF, qL, yF, yL, xp, xn);
    for (int g = 0; g < G; ++g) { xprev.p[g][col0 + c] = xp[g]; xnext.p[g][col0 + c] = xn[g]; }
  }
}

__global__ void dz_fill_plane_kernel(double* __restrict__ plane, long n, double v) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) plane[q] = v;
}

// ---- the reference-order columns in the distributed layout ----------------------------------------------------
// sel[q] = global column index of selected column q (sorted); owner of column c = c / ncol_own; selected columns
// [own0[r], own0[r+1]) belong to rank r.  Gather: rank `rank` writes its n3l levels of EVERY selected column into the
// owner's dense (nsel_own x nz) matrix G_owner[(q - own0) + nsel_own * (k0 + k)].
struct DzOwn { int own0[FB_DZ_MAXG + 1]; };
__global__ void dz_gather_sel_kernel(int nsel, int n3l, int k0, long ncol, long ncol_own, const int* __restrict__ sel,
                                     const double* __restrict__ W, DzOwn ow, DzPeers gath) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nsel * n3l) return;
  const int q = (int)(idx % nsel), k = (int)(idx / nsel);
  const long col = sel[q];
  const int r = (int)(col / ncol_own);
  const int nown = ow.own0[r + 1] - ow.own0[r];
  gath.p[r][(q - ow.own0[r]) + (long)nown * (k0 + k)] = W[col + ncol * (long)k];
}
// owner: F[nsel_own][nz] (thomas_ref.cuh) -> every rank's override buffer OVR_g[(own0 + q) * n3l + k]
__global__ void dz_push_sel_kernel(int nsel_own, int own0, int nz, int n3l, const double* __restrict__ F, DzPeers ovr) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nsel_own * nz) return;
  const int q = (int)(idx / nz), k = (int)(idx - (long)q * nz);
  const int g = k / n3l;
  ovr.p[g][(long)(own0 + q) * n3l + (k - g * n3l)] = F[idx];
}
// every rank: override buffer -> its slab
__global__ void dz_apply_sel_kernel(int nsel, int n3l, long ncol, const int* __restrict__ sel, const double* __restrict__ OVR,
                                    double* __restrict__ W) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nsel * n3l) return;
  const int q = (int)(idx / n3l), k = (int)(idx - (long)q * n3l);
  W[(long)sel[q] + ncol * (long)k] = OVR[idx];
}

}  // namespace fb
#endif
