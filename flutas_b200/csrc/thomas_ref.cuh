// thomas_ref.cuh -- reference-order z solve for the few ILL-CONDITIONED (kx,ky) columns.
//
// Why: the partition + PCR kernels (thomas_uni / thomas_reg / thomas_tile) and the reference's sequential
// dgtsv_homebrewed (src/solver_cpu.f90:187-223) are both backward stable, but they round DIFFERENTLY, and for a
// column with a small |lambdaxy| the system b + lambda is nearly singular: cond ~ 4 max(a) / |lambda| (4e6 for the
// gravest modes of a 1024-level grid in lz = 1).  Two stable eliminations then differ by ~cond * eps / sqrt(nz)
// -- measured 2-3e-12 of max|p| on the 1024^3 channel grid, above the 1e-12 parity bar, while either result is
// 1e-11 away from the exactly rounded solution (tests/test_thomas_ref.py).  The only way to agree with the reference
// beyond its own conditioning is to round like it where the conditioning lives.
//
// Where the conditioning lives: in the LU factors.  z_l = 1/(bb_l - a_l d_{l-1}), d_l = c_l z_l and the last pivot
// bb_n - a_n d_{n-1} (catastrophic cancellation down to O(lambda n)) depend only on a, b, c, lambda -- NOT on the
// right-hand side.  So, once per plan and per selected column, the factors are computed sequentially in exactly the
// reference's operation order (no FMA contraction, IEEE division): `ref_factor`.  Per solve, only the two right-hand-side
// recurrences remain,
//     y_l = (p_l - a_l y_{l-1}) z_l            x_l = y_l - d_l x_{l+1},
// whose rounding errors act like relative perturbations of the right-hand side (~eps sqrt(nz), NOT amplified by cond), so
// they may run in parallel: a warp owns a column, each lane sweeps a segment of ~nz/32 levels, the segment inflows come
// from a composition of the per-segment affine maps.  Periodic z follows gaussel_periodic (src/solver_cpu.f90:147-185):
// the (n-1)-row solve for p1 as above; p2 and the closure denominator are right-hand-side independent and precomputed.
//
// Selection: |lambda| < 4 max(|a|,|c|) * tol (tol = 1e-5: 132 of 1 M columns at 1024^3, 6-17 on the other BASELINE
// grids).  The selected columns are solved into a side buffer BEFORE the main kernel (which overwrites the work
// array in place) and scattered over its result afterwards: two tiny launches, no change to the main kernels.
// The singular column (lambda = 0, all-Neumann/periodic problem) keeps the gauge of the main kernels, x(nz) = 0.
//
// Host-compilable core (tests/emulate), like the other z kernels.
#pragma once
#include "thomas_tile.cuh"

namespace fb {

// the reference's expressions, one rounding per operation (gfortran without contraction; the oracle is built with
// -ffp-contract=off).  Host build (tests/emulate): plain operators, compiled with -ffp-contract=off.
#if defined(__CUDA_ARCH__)
#define FB_XMUL(a, b) __dmul_rn((a), (b))
#define FB_XADD(a, b) __dadd_rn((a), (b))
#define FB_XSUB(a, b) __dsub_rn((a), (b))
#define FB_XDIV(a, b) __ddiv_rn((a), (b))
#else
#define FB_XMUL(a, b) ((a) * (b))
#define FB_XADD(a, b) ((a) + (b))
#define FB_XSUB(a, b) ((a) - (b))
#define FB_XDIV(a, b) ((a) / (b))
#endif

// Device-visible tables of the selected columns.  m = rows of the system dgtsv_homebrewed factorises (nz, or nz - 1
// when z is periodic); z, d: [nsel][nz] (entries 0..m-2 used); p2: [nsel][nz] (periodic only, entries 0..m-1).
struct RefTables {
  int nsel, nz, m, periodic;
  const double *a, *b, *c;      // a(1:n), b(1:n), c(1:n) exactly as initsolver leaves them (a(1), c(n) NOT zeroed)
  const int* col;               // [nsel] column index in the (ncol x nz) work array
  const double* lam;            // [nsel]
  const unsigned char* pin;     // [nsel] 1: singular column, x(nz) = 0
  double *z, *d, *piv, *p2, *den;
};

// ---- once per plan: LU factors of one column in the reference's order (dgtsv_homebrewed, solver_cpu.f90:201-208) ----
FB_HD void ref_factor(const RefTables& R, int q) {
  const int m = R.m, n = R.nz;
  const double lam = R.lam[q];
  double* z = R.z + (size_t)q * n;
  double* d = R.d + (size_t)q * n;
  double zz = FB_XDIV(1.0, FB_XADD(R.b[0], lam));                       // z = 1/b(1)       (bb = b + lambdaxy, :135,166)
  double dd = FB_XMUL(R.c[0], zz);                                      // d(1) = c(1)*z
  z[0] = zz; d[0] = dd;
  for (int l = 1; l < m - 1; ++l) {
    zz = FB_XDIV(1.0, FB_XSUB(FB_XADD(R.b[l], lam), FB_XMUL(R.a[l], dd)));   // z = 1/(b(l)-a(l)*d(l-1))
    dd = FB_XMUL(R.c[l], zz);                                           // d(l) = c(l)*z
    z[l] = zz; d[l] = dd;
  }
  const double piv = FB_XSUB(FB_XADD(R.b[m - 1], lam), FB_XMUL(R.a[m - 1], dd));   // z = b(n)-a(n)*d(n-1)
  R.piv[q] = piv;
  if (!R.periodic) return;
  // gaussel_periodic (:168-176): p2 = (-a(1), 0, ..., 0, -c(n-1)) solved with the same factors, and the closure denominator
  double* p2 = R.p2 + (size_t)q * n;
  double y = FB_XMUL(-R.a[0], z[0]);                                    // p(1) = p(1)*z
  p2[0] = y;
  for (int l = 1; l < m - 1; ++l) {
    y = FB_XMUL(FB_XSUB(0.0, FB_XMUL(R.a[l], y)), z[l]);                // p(l) = (p(l)-a(l)*p(l-1))*z
    p2[l] = y;
  }
  double x = (piv != 0.0) ? FB_XDIV(FB_XSUB(-R.c[m - 1], FB_XMUL(R.a[m - 1], y)), piv) : 0.0;
  p2[m - 1] = x;
  for (int l = m - 2; l >= 0; --l) {                                    // p(l) = p(l) - d(l)*p(l+1)
    x = FB_XSUB(p2[l], FB_XMUL(d[l], x));
    p2[l] = x;
  }
  // bb(n) + c(n)*p2(1) + a(n)*p2(n-1)
  R.den[q] = FB_XADD(FB_XADD(FB_XADD(R.b[n - 1], lam), FB_XMUL(R.c[n - 1], p2[0])), FB_XMUL(R.a[n - 1], p2[m - 1]));
}

// ---- per solve: one warp per column ---------------------------------------------------------------------------
// Staging layout of a column in shared memory: level l of segment t = l / Lc sits at t*LP + (l - t*Lc), LP = Lc | 1
// (odd pitch: the 32 lanes of a sweep hit different banks).  Arrays: p (right-hand side -> y -> x), z, d.
struct RefShape {
  int Lc, LP;                   // levels per lane, padded pitch
  FB_HD explicit RefShape(int nz) : Lc((nz + 31) / 32), LP(((nz + 31) / 32) | 1) {}
  FB_HD int at(int l) const { const int t = l / Lc; return t * LP + (l - t * Lc); }
  FB_HD int doubles() const { return 32 * LP; }
};

// rows [l0, l1) of lane `lane` among rows 0..last-1
FB_HD void ref_segment(const RefShape& S, int lane, int last, int& l0, int& l1) {
  l0 = lane * S.Lc; l1 = l0 + S.Lc;
  if (l0 > last) l0 = last;
  if (l1 > last) l1 = last;
}

// forward recurrence over rows 0..m-2, pass 1: zero inflow; affine map (PA, PB) of the segment
FB_HD void ref_fwd_local(const RefTables& R, const RefShape& S, const double* p, const double* z, int lane, double* PA, double* PB) {
  int l0, l1;
  ref_segment(S, lane, R.m - 1, l0, l1);
  double yy = 0.0, pa = 1.0;
  const int o = lane * S.LP - l0;
  for (int l = l0; l < l1; ++l) {
    if (l == 0) { yy = p[o + l] * z[o + l]; pa = 0.0; }
    else { const double al = R.a[l], zl = z[o + l]; yy = (p[o + l] - al * yy) * zl; pa = -(al * zl) * pa; }
  }
  PA[lane] = pa; PB[lane] = yy;
}
// inflow of every segment (one lane): IN[t] = value entering segment t; returns the value leaving the last one
FB_HD double ref_chain_up(const double* PA, const double* PB, double* IN, double start) {
  double cur = start;
  for (int t = 0; t < 32; ++t) { IN[t] = cur; cur = PA[t] * cur + PB[t]; }
  return cur;
}
FB_HD double ref_chain_down(const double* PA, const double* PB, double* IN, double start) {
  double cur = start;
  for (int t = 31; t >= 0; --t) { IN[t] = cur; cur = PA[t] * cur + PB[t]; }
  return cur;
}
// pass 2: the reference's expressions with the true inflow; p <- y
FB_HD void ref_fwd_final(const RefTables& R, const RefShape& S, double* p, const double* z, int lane, double yin) {
  int l0, l1;
  ref_segment(S, lane, R.m - 1, l0, l1);
  double yy = yin;
  const int o = lane * S.LP - l0;
  for (int l = l0; l < l1; ++l) {
    if (l == 0) yy = FB_XMUL(p[o + l], z[o + l]);                                          // p(1) = p(1)*z
    else yy = FB_XMUL(FB_XSUB(p[o + l], FB_XMUL(R.a[l], yy)), z[o + l]);                   // p(l) = (p(l)-a(l)*p(l-1))*z
    p[o + l] = yy;
  }
}
// last row of the m-row system (solver_cpu.f90:209-214); returns x(m)
FB_HD double ref_last_row(const RefTables& R, const RefShape& S, double* p, int q, bool pin_here) {
  const int m = R.m;
  const double piv = R.piv[q];
  const double ym = (m >= 2) ? p[S.at(m - 2)] : 0.0;
  double x = 0.0;
  if (!pin_here && piv != 0.0) x = FB_XDIV(FB_XSUB(p[S.at(m - 1)], FB_XMUL(R.a[m - 1], ym)), piv);
  p[S.at(m - 1)] = x;
  return x;
}
// backward recurrence x_l = y_l - d_l x_{l+1} over rows m-2..0
FB_HD void ref_bwd_local(const RefTables& R, const RefShape& S, const double* p, const double* d, int lane, double* PA, double* PB) {
  int l0, l1;
  ref_segment(S, lane, R.m - 1, l0, l1);
  double xx = 0.0, pa = 1.0;
  const int o = lane * S.LP - l0;
  for (int l = l1 - 1; l >= l0; --l) { const double dl = d[o + l]; xx = p[o + l] - dl * xx; pa = -dl * pa; }
  PA[lane] = pa; PB[lane] = xx;
}
FB_HD void ref_bwd_final(const RefTables& R, const RefShape& S, double* p, const double* d, int lane, double xin) {
  int l0, l1;
  ref_segment(S, lane, R.m - 1, l0, l1);
  double xx = xin;
  const int o = lane * S.LP - l0;
  for (int l = l1 - 1; l >= l0; --l) { xx = FB_XSUB(p[o + l], FB_XMUL(d[o + l], xx)); p[o + l] = xx; }   // p(l) = p(l) - d(l)*p(l+1)
}
// periodic closure (solver_cpu.f90:175-177): returns p(n); p holds p1(1:n-1) and the untouched right-hand side p(n)
FB_HD double ref_closure(const RefTables& R, const RefShape& S, const double* p, int q, bool pin) {
  if (pin) return 0.0;
  const int n = R.nz, m = R.m;
  const double num = FB_XSUB(FB_XSUB(p[S.at(n - 1)], FB_XMUL(R.c[n - 1], p[S.at(0)])), FB_XMUL(R.a[n - 1], p[S.at(m - 1)]));
  return FB_XDIV(num, R.den[q]);
}

}  // namespace fb

#if defined(__CUDACC__)
#include "geom.cuh"

namespace fb {

__global__ void ref_factor_kernel(RefTables R) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < R.nsel) ref_factor(R, q);
}

// Solves the selected columns of W (ncol x nz, column `col`, level k at W[col + ncol*k]) into F[nsel][nz].
// blockDim.x = 32 * warps; dynamic shared memory = warps * (3 * 32 LP + 96) doubles.
__global__ void ref_solve_kernel(RefTables R, long ncol, const double* __restrict__ W, double* __restrict__ F) {
  extern __shared__ double smem[];
  const RefShape S(R.nz);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const size_t per = 3 * (size_t)S.doubles() + 96;
  double* p = smem + warp * per;
  double* z = p + S.doubles();
  double* d = z + S.doubles();
  double* PA = d + S.doubles();
  double* PB = PA + 32;
  double* IN = PB + 32;
  const int n = R.nz, m = R.m;
  for (int q = blockIdx.x * nwarp + warp; q < R.nsel; q += gridDim.x * nwarp) {
    const long col = R.col[q];
    const bool pin = R.pin[q] != 0;
    const double* zq = R.z + (size_t)q * n;
    const double* dq = R.d + (size_t)q * n;
    for (int l = lane; l < n; l += 32) {
      const int o = S.at(l);
      p[o] = __ldcs(W + col + ncol * (long)l);
      if (l < m - 1) { z[o] = __ldg(zq + l); d[o] = __ldg(dq + l); }
    }
    __syncwarp();
    ref_fwd_local(R, S, p, z, lane, PA, PB);
    __syncwarp();
    if (lane == 0) ref_chain_up(PA, PB, IN, 0.0);
    __syncwarp();
    ref_fwd_final(R, S, p, z, lane, IN[lane]);
    __syncwarp();
    double xm = 0.0;
    if (lane == 0) { xm = ref_last_row(R, S, p, q, pin && !R.periodic); PA[0] = xm; }
    __syncwarp();
    xm = PA[0];
    __syncwarp();
    ref_bwd_local(R, S, p, d, lane, PA, PB);
    __syncwarp();
    if (lane == 0) ref_chain_down(PA, PB, IN, xm);
    __syncwarp();
    ref_bwd_final(R, S, p, d, lane, IN[lane]);
    __syncwarp();
    double* Fq = F + (size_t)q * n;
    if (R.periodic) {
      const double pn = ref_closure(R, S, p, q, pin);                    // every lane evaluates the same expression
      const double* p2 = R.p2 + (size_t)q * n;
      for (int l = lane; l < m; l += 32) Fq[l] = FB_XADD(p[S.at(l)], FB_XMUL(__ldg(p2 + l), pn));   // p1 + p2*p(n)
      if (lane == 0) Fq[n - 1] = pn;
    } else {
      for (int l = lane; l < n; l += 32) Fq[l] = p[S.at(l)];
    }
    __syncwarp();
  }
}

// F[nsel][nz] -> the z stage's output geometry (ColGeom: the work array itself, or the peers' receive buffers)
__global__ void ref_scatter_kernel(RefTables R, long ncol, const double* __restrict__ F, ColGeom og) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)R.nsel * R.nz) return;
  const int q = (int)(idx / R.nz), k = (int)(idx - (long)q * R.nz);
  const int r = k / og.n3l;
  og.ptr[r][og.koff + R.col[q] + ncol * (long)(k - r * og.n3l)] = F[idx];
}

}  // namespace fb
#endif
