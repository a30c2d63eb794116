// thomas_hier.cuh -- hierarchical solve of the reduced (separator) system for a z column spread over G CTAs.
//
// STATUS: numerical core only (host-compilable, exercised by tests/emulate + tests/test_tile_core.py).  No kernel uses
// it yet; it is the groundwork for 16-column tiles on stretched grids with nz = 1024.
//
// Why: thomas_reg_kernel with CL = 2 (two CTAs of a cluster split the levels of a column) solves the WHOLE reduced system
// redundantly in both CTAs and ships seven doubles per segment through DSMEM -- measured slower than 8-column tiles.
// Here each CTA keeps only ITS m = S/G unit-diagonal reduced rows  A_s X_{s-1} + X_s + C_s X_{s+1} = R_s  and
//   1. runs PCR on its local block with THREE right-hand sides (R, the left coupling A_first e_first, the right coupling
//      C_last e_last); afterwards every local unknown is  X_s = y_s - p_s X_left - q_s X_right  with X_left / X_right the
//      neighbouring groups' last / first unknown;
//   2. publishes six doubles per column, (y,p,q) of its first and last row -- all that ever crosses the SM boundary;
//   3. solves the 2G x 2G interface system for (first, last) of every group (redundantly, it is tiny) and substitutes.
// Periodic z couples group 0 to group G-1 through the same formulas; a non-periodic system simply has A = 0 in the very
// first row and C = 0 in the very last one.
#pragma once
#include "thomas_reg.cuh"

namespace fb {

enum { FB_HIER_MAXG = 4 };

template <int TI>
struct ThomasHier {
  // one local PCR step with stride h on row s (global index) of group g = s / m; src/dst hold 5 arrays [A | C | y | p | q]
  // of S*TI doubles each.  Neighbours outside the group do not exist for the local problem (their coupling is carried by
  // p and q instead), so they are clamped to the row itself with the coupling coefficient forced to zero.
  static FB_HD bool pcr3_step(const double* src, double* dst, int S, int m, int lane, int s, int h) {
    const int st = S * TI, o = s * TI + lane;
    const int g0 = (s / m) * m, g1 = g0 + m - 1;
    const int sm = s - h, sp = s + h;
    const bool hm = sm >= g0, hp = sp <= g1;
    const int qm = (hm ? sm : s) * TI + lane, qp = (hp ? sp : s) * TI + lane;
    const double A = hm ? src[o] : 0.0, C = hp ? src[st + o] : 0.0;
    const double Am = src[qm], Cm = src[st + qm], Ap = src[qp], Cp = src[st + qp];
    const double inv = fb_rcp(1.0 - A * Cm - C * Ap);
    const double An = -A * Am * inv, Cn = -C * Cp * inv;
    dst[o] = hm ? An : 0.0;
    dst[st + o] = hp ? Cn : 0.0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 2; r < 5; ++r) dst[r * st + o] = (src[r * st + o] - A * src[r * st + qm] - C * src[r * st + qp]) * inv;
    const double tiny = 8.6736173798840355e-19;                            // 2^-60, as ThomasReg::pcr_step
    return fabs(dst[o]) > tiny || fabs(dst[st + o]) > tiny;
  }

  // initial 5-array state of row s from the normalised reduced row (A, C, R) = pcr[0..2]
  static FB_HD void init_row(const double* pcr, double* w, int S, int m, int lane, int s) {
    const int st = S * TI, o = s * TI + lane;
    const int g0 = (s / m) * m, g1 = g0 + m - 1;
    const double A = pcr[o], C = pcr[st + o];
    w[o] = (s == g0) ? 0.0 : A;                                            // couplings leaving the group move to p / q
    w[st + o] = (s == g1) ? 0.0 : C;
    w[2 * st + o] = pcr[2 * st + o];
    w[3 * st + o] = (s == g0) ? A : 0.0;
    w[4 * st + o] = (s == g1) ? C : 0.0;
  }

  // interface system for one column: unknowns u[2g] = first, u[2g+1] = last unknown of group g.
  //   first_g + p_f(g) last_{g-1} + q_f(g) first_{g+1} = y_f(g),   last_g + p_l(g) last_{g-1} + q_l(g) first_{g+1} = y_l(g)
  // (indices modulo G; the couplings that do not exist in a non-periodic system are zero in p / q already).
  // Dense elimination with partial pivoting on at most 8 x 8.
  static FB_HD void interface_solve(const double* w, int S, int G, int lane, double* u) {
    const int st = S * TI, m = S / G, n = 2 * G;
    double Mx[2 * FB_HIER_MAXG][2 * FB_HIER_MAXG + 1];
    for (int r = 0; r < n; ++r) for (int c = 0; c <= n; ++c) Mx[r][c] = 0.0;
    for (int g = 0; g < G; ++g) {
      const int gl = (g + G - 1) % G, gr = (g + 1) % G;
      for (int e = 0; e < 2; ++e) {                                        // e = 0: first row of the group, 1: last row
        const int s = g * m + (e ? m - 1 : 0), o = s * TI + lane, r = 2 * g + e;
        Mx[r][r] += 1.0;
        Mx[r][2 * gl + 1] += w[3 * st + o];                                // p: coupling to the left group's last unknown
        Mx[r][2 * gr] += w[4 * st + o];                                    // q: coupling to the right group's first unknown
        Mx[r][n] = w[2 * st + o];
      }
    }
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int r = c + 1; r < n; ++r) if (fabs(Mx[r][c]) > fabs(Mx[piv][c])) piv = r;
      if (piv != c) for (int k = 0; k <= n; ++k) { const double t = Mx[c][k]; Mx[c][k] = Mx[piv][k]; Mx[piv][k] = t; }
      const double inv = 1.0 / Mx[c][c];
      for (int r = c + 1; r < n; ++r) {
        const double f = Mx[r][c] * inv;
        if (f != 0.0) for (int k = c; k <= n; ++k) Mx[r][k] -= f * Mx[c][k];
      }
    }
    for (int r = n - 1; r >= 0; --r) {
      double acc = Mx[r][n];
      for (int k = r + 1; k < n; ++k) acc -= Mx[r][k] * u[k];
      u[r] = acc / Mx[r][r];
    }
  }

  // X_s = y_s - p_s X_left - q_s X_right with the interface values of the neighbouring groups
  static FB_HD void substitute(const double* w, const double* u, double* X, int S, int G, int lane, int s) {
    const int st = S * TI, m = S / G, o = s * TI + lane, g = s / m;
    const double xl = u[2 * ((g + G - 1) % G) + 1], xr = u[2 * ((g + 1) % G)];
    X[o] = w[2 * st + o] - w[3 * st + o] * xl - w[4 * st + o] * xr;
  }
};

}  // namespace fb
