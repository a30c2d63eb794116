/*
 * flutas_oracle.c -- CPU restatement of the FluTAS constant-coefficient pressure-Poisson path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (libflutas_b200.so)
 * never links, loads or calls anything in this file.
 *
 * What it restates (file:line relative to the FluTAS tree):
 *   oracle_fillps        src/fillps.f90:16-69      (with _CONSTANT_COEFFS_POISSON, always on: apps/<APP>/app.<APP>:2)
 *   oracle_updt_rhs_b    src/bound.f90:829-944     (single rank: every face is a domain boundary)
 *   oracle_eigenvalues   src/initsolver.f90:122-186 (non-_OPENACC branch)
 *   oracle_tridmatrix    src/initsolver.f90:188-246
 *   oracle_find_fft      src/fft.f90:233-291
 *   oracle_normfft       src/fft.f90:71,87,125,150
 *   oracle_solver_cpu    src/solver_cpu.f90:20-115 (one rank: the 2DECOMP transposes are identities)
 *   gaussel / gaussel_periodic / dgtsv_homebrewed  src/solver_cpu.f90:117-223 (exact operation order; by default swept over eight
 *                        adjacent columns at a time -- bit-identical to the column-at-a-time transcription, which
 *                        oracle_set_definition_path(1) selects and tests/test_oracle.py compares bit for bit)
 *   oracle_correc        src/correc.f90:16-81      (_CONSTANT_COEFFS_POISSON branch)
 *   oracle_pres_sp_src   src/source.f90:311-346    oracle_pres_tw_src  src/source.f90:247-309 (constant-coefficient branch)
 *   oracle_pold_update   src/apps/single_phase/main__single_phase.f90:693-699, 734-740
 *   oracle_chkdiv        src/chkdiv.f90:18-69
 *   oracle_initgrid      src/initgrid.f90:17-118   (two-end tanh clustering)
 *
 * Third-party arithmetic: the reference calls FFTW 3 (not vendored; docs pin fftw-3.3.10,
 * getting_started/REQ.md:9-22) through fftw_plan_guru_r2r / dfftw_execute_r2r
 * (src/fft.f90:85-86,123-124,188-190).  FFTW is absent here, so the r2r kinds are restated from
 * FFTW's published definitions (unnormalised):
 *   R2HC    Y_k = sum_j x_j e^{-2 pi i jk/n}, stored r0..r_{n/2}, i_{(n+1)/2-1}..i_1
 *   HC2R    inverse of the above without 1/n
 *   REDFT10 Y_k = 2 sum_j x_j cos(pi (j+1/2) k / n)
 *   REDFT01 Y_k = x_0 + 2 sum_{j>=1} x_j cos(pi j (k+1/2) / n)
 *   RODFT10 Y_k = 2 sum_j x_j sin(pi (j+1/2)(k+1) / n)
 *   RODFT01 Y_k = (-1)^k x_{n-1} + 2 sum_{j<n-1} x_j sin(pi (j+1)(k+1/2) / n)
 *   REDFT11 Y_k = 2 sum_j x_j cos(pi (j+1/2)(k+1/2) / n)
 *   RODFT11 Y_k = 2 sum_j x_j sin(pi (j+1/2)(k+1/2) / n)
 * each evaluated through ONE generic complex FFT (length n for R2HC/HC2R, 2n zero-padded for
 * the DCT/DST kinds) so the code is a direct transcription of the formulas.
 *
 * PARITY STATUS: "parity unpinned" against the reference's own vectors -- the reference holds
 * no solver-level golden vectors and cannot be compiled here (no Fortran
 * compiler, MPI or FFTW).  The oracle is pinned instead to (i) scipy.fft/pocketfft, whose
 * rfft/dct/dst definitions equal FFTW's, (ii) an O(n^2) long-double evaluation of the
 * definitions above, (iii) the residual of the discrete 7-point operator (tests/test_oracle.py).
 *
 * Arrays are Fortran column-major, passed as flat pointers.  p has a 1-cell halo:
 *   p(i,j,k), i=0..n1+1  ->  p[i + (n1+2)*(j + (n2+2)*k)]
 * u,v,w have an nh_u-cell halo (lower bound 1-nh_u); dzci/dzfi have lower bound 1-nh_d.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FFTW_R2HC 0
#define FFTW_HC2R 1
#define FFTW_REDFT00 3
#define FFTW_REDFT01 4
#define FFTW_REDFT10 5
#define FFTW_REDFT11 6
#define FFTW_RODFT00 7
#define FFTW_RODFT01 8
#define FFTW_RODFT10 9
#define FFTW_RODFT11 10

typedef struct { double re, im; } cplx;

static const long double PI_L = 3.14159265358979323846264338327950288L;

/* ------------------------------------------------------------------ generic complex FFT */

typedef struct {
  int n;       /* complex length */
  cplx *tw;    /* tw[k] = exp(-2 pi i k / n) */
} cfft_plan;

static cfft_plan *cfft_create(int n) {
  cfft_plan *pl = (cfft_plan *)malloc(sizeof(cfft_plan));
  pl->n = n;
  pl->tw = (cplx *)malloc(sizeof(cplx) * (size_t)n);
  for (int k = 0; k < n; ++k) {
    long double ang = -2.0L * PI_L * (long double)k / (long double)n;
    pl->tw[k].re = (double)cosl(ang);
    pl->tw[k].im = (double)sinl(ang);
  }
  return pl;
}
static void cfft_destroy(cfft_plan *pl) { if (pl) { free(pl->tw); free(pl); } }

static int smallest_factor(int n) {
  if (n % 4 == 0) return 4;
  if (n % 2 == 0) return 2;
  for (int p = 3; p * p <= n; p += 2) if (n % p == 0) return p;
  return n;
}

/* out[0..n) = sum_j in[j*is] w^{jk}, w = exp(sign * 2 pi i / n); tws = N_top / n */
static void cfft_rec(const cfft_plan *pl, int n, int is, const cplx *in, cplx *out, int tws, int sign) {
  if (n == 1) { out[0] = in[0]; return; }
  const int p = smallest_factor(n), m = n / p, N = pl->n;
  for (int r = 0; r < p; ++r) cfft_rec(pl, m, is * p, in + (size_t)r * is, out + (size_t)r * m, tws * p, sign);
  cplx t[64];
  cplx *tt = t, *heap = NULL;
  if (p > 64) { heap = (cplx *)malloc(sizeof(cplx) * (size_t)p); tt = heap; }
  for (int k = 0; k < m; ++k) {
    for (int r = 0; r < p; ++r) {                 /* twiddle: w_n^{r k} */
      cplx w = pl->tw[(int)(((long long)r * k * tws) % N)];
      if (sign > 0) w.im = -w.im;
      cplx a = out[(size_t)r * m + k];
      tt[r].re = a.re * w.re - a.im * w.im;
      tt[r].im = a.re * w.im + a.im * w.re;
    }
    if (p == 2) {
      out[k].re = tt[0].re + tt[1].re; out[k].im = tt[0].im + tt[1].im;
      out[k + m].re = tt[0].re - tt[1].re; out[k + m].im = tt[0].im - tt[1].im;
    } else if (p == 4) {
      cplx s02 = { tt[0].re + tt[2].re, tt[0].im + tt[2].im }, d02 = { tt[0].re - tt[2].re, tt[0].im - tt[2].im };
      cplx s13 = { tt[1].re + tt[3].re, tt[1].im + tt[3].im }, d13 = { tt[1].re - tt[3].re, tt[1].im - tt[3].im };
      /* multiply d13 by (-i) for sign<0, (+i) for sign>0 */
      cplx jd = (sign < 0) ? (cplx){ d13.im, -d13.re } : (cplx){ -d13.im, d13.re };
      out[k].re = s02.re + s13.re;         out[k].im = s02.im + s13.im;
      out[k + m].re = d02.re + jd.re;      out[k + m].im = d02.im + jd.im;
      out[k + 2 * m].re = s02.re - s13.re; out[k + 2 * m].im = s02.im - s13.im;
      out[k + 3 * m].re = d02.re - jd.re;  out[k + 3 * m].im = d02.im - jd.im;
    } else {
      for (int q = 0; q < p; ++q) {               /* w_p^{r q} = tw[(r q m tws) mod N] */
        double sr = 0.0, si = 0.0;
        for (int r = 0; r < p; ++r) {
          cplx w = pl->tw[(int)(((long long)r * q % p) * m * tws % N)];
          if (sign > 0) w.im = -w.im;
          sr += tt[r].re * w.re - tt[r].im * w.im;
          si += tt[r].re * w.im + tt[r].im * w.re;
        }
        out[k + (size_t)q * m].re = sr; out[k + (size_t)q * m].im = si;
      }
    }
  }
  if (heap) free(heap);
}

static void cfft_exec(const cfft_plan *pl, const cplx *in, cplx *out, int sign) {
  cfft_rec(pl, pl->n, 1, in, out, 1, sign);
}

/* ------------------------------------------------------------------ r2r plans */

typedef struct {
  int n, kind;
  cfft_plan *cf;      /* length n (R2HC/HC2R) or 2n (DCT/DST kinds) */
  cplx *ph;           /* phase tables, see r2r_exec */
  cplx *ph2;
  /* fast path (even n): half-length complex FFT + tables, see r2r_exec_fast */
  cfft_plan *ch;      /* length n/2 */
  cplx *wn;           /* wn[k] = exp(-2 pi i k / n),        k = 0..n/2 */
  cplx *wq;           /* wq[k] = exp(-i pi k / (2n)),       k = 0..n/2 */
  cplx *w4a;          /* w4a[m] = exp(-i pi (4m+1)/(4n)),   m = 0..n/2-1 */
  cplx *w4b;          /* w4b[k] = exp(-i pi k / n),         k = 0..n/2-1 */
} r2r_plan;

/* 1: every transform goes through the definition-transcribing path (r2r_exec); 0 (default): even lengths use the fast path,
 * which tests/test_oracle.py holds against the definition path, scipy/pocketfft and the long-double O(n^2) sums */
static int g_force_definition_path = 0;
void oracle_set_definition_path(int on) { g_force_definition_path = on; }

static cplx unit_phase(long double num, long double den) {      /* exp(-i pi num/den) */
  cplx w; long double ang = -PI_L * num / den;
  w.re = (double)cosl(ang); w.im = (double)sinl(ang);
  return w;
}

void *oracle_r2r_create(int n, int kind) {
  r2r_plan *pl = (r2r_plan *)calloc(1, sizeof(r2r_plan));
  pl->n = n; pl->kind = kind;
  switch (kind) {
    case FFTW_R2HC: case FFTW_HC2R:
      pl->cf = cfft_create(n); break;
    case FFTW_REDFT10: case FFTW_RODFT10: case FFTW_REDFT01: case FFTW_RODFT01:
    case FFTW_REDFT11: case FFTW_RODFT11:
      pl->cf = cfft_create(2 * n);
      /* ph[k] = exp(-i pi k / (2n)), k = 0..2n ; ph2[k] = exp(-i pi (2k+1)/(4n)), k=0..n-1 */
      pl->ph = (cplx *)malloc(sizeof(cplx) * (size_t)(2 * n + 1));
      for (int k = 0; k <= 2 * n; ++k) {
        long double ang = -PI_L * (long double)k / (2.0L * n);
        pl->ph[k].re = (double)cosl(ang); pl->ph[k].im = (double)sinl(ang);
      }
      pl->ph2 = (cplx *)malloc(sizeof(cplx) * (size_t)n);
      for (int k = 0; k < n; ++k) {
        long double ang = -PI_L * (long double)(2 * k + 1) / (4.0L * n);
        pl->ph2[k].re = (double)cosl(ang); pl->ph2[k].im = (double)sinl(ang);
      }
      break;
    default:
      free(pl); return NULL;   /* face-centred kinds REDFT00/RODFT00: dead code in FluTAS */
  }
  if (n % 2 == 0 && n >= 4) {
    const int M = n / 2;
    pl->ch = cfft_create(M);
    pl->wn = (cplx *)malloc(sizeof(cplx) * (size_t)(M + 1));
    pl->wq = (cplx *)malloc(sizeof(cplx) * (size_t)(M + 1));
    pl->w4a = (cplx *)malloc(sizeof(cplx) * (size_t)M);
    pl->w4b = (cplx *)malloc(sizeof(cplx) * (size_t)M);
    for (int k = 0; k <= M; ++k) { pl->wn[k] = unit_phase(2.0L * k, (long double)n); pl->wq[k] = unit_phase((long double)k, 2.0L * n); }
    for (int k = 0; k < M; ++k) { pl->w4a[k] = unit_phase(4.0L * k + 1.0L, 4.0L * n); pl->w4b[k] = unit_phase((long double)k, (long double)n); }
  }
  return pl;
}

void oracle_r2r_destroy(void *h) {
  r2r_plan *pl = (r2r_plan *)h;
  if (!pl) return;
  cfft_destroy(pl->cf); free(pl->ph); free(pl->ph2);
  cfft_destroy(pl->ch); free(pl->wn); free(pl->wq); free(pl->w4a); free(pl->w4b); free(pl);
}

/* transform one line x[0..n) (contiguous) in place; wa, wb: scratch of 2n cplx each */
static void r2r_exec(const r2r_plan *pl, double *x, cplx *wa, cplx *wb) {
  const int n = pl->n;
  switch (pl->kind) {
    case FFTW_R2HC: {
      for (int j = 0; j < n; ++j) { wa[j].re = x[j]; wa[j].im = 0.0; }
      cfft_exec(pl->cf, wa, wb, -1);
      for (int k = 0; k <= n / 2; ++k) x[k] = wb[k].re;
      for (int k = 1; k < (n + 1) / 2; ++k) x[n - k] = wb[k].im;
    } break;
    case FFTW_HC2R: {
      wa[0].re = x[0]; wa[0].im = 0.0;
      for (int k = 1; k < (n + 1) / 2; ++k) {
        wa[k].re = x[k]; wa[k].im = x[n - k];
        wa[n - k].re = x[k]; wa[n - k].im = -x[n - k];
      }
      if (n % 2 == 0) { wa[n / 2].re = x[n / 2]; wa[n / 2].im = 0.0; }
      cfft_exec(pl->cf, wa, wb, +1);
      for (int j = 0; j < n; ++j) x[j] = wb[j].re;
    } break;
    case FFTW_REDFT10: case FFTW_RODFT10: {
      /* F_k = sum_j x_j e^{-2 pi i jk/(2n)} ; REDFT10: Y_k = Re(2 e^{-i pi k/(2n)} F_k)
         RODFT10: Y_k = -Im(2 e^{-i pi (k+1)/(2n)} F_{k+1}) */
      for (int j = 0; j < n; ++j) { wa[j].re = x[j]; wa[j].im = 0.0; wa[n + j].re = 0.0; wa[n + j].im = 0.0; }
      cfft_exec(pl->cf, wa, wb, -1);
      if (pl->kind == FFTW_REDFT10) {
        for (int k = 0; k < n; ++k) x[k] = 2.0 * (pl->ph[k].re * wb[k].re - pl->ph[k].im * wb[k].im);
      } else {
        for (int k = 0; k < n; ++k) {
          cplx f = wb[(k + 1) % (2 * n)], p = pl->ph[k + 1];
          x[k] = -2.0 * (p.re * f.im + p.im * f.re);
        }
      }
    } break;
    case FFTW_REDFT01: {
      /* Y_k = Re( sum_j c_j x_j e^{+i pi j/(2n)} e^{+2 pi i jk/(2n)} ), c_0 = 1, c_j = 2 */
      for (int j = 0; j < n; ++j) {
        double c = (j == 0) ? 1.0 : 2.0;
        wa[j].re = c * x[j] * pl->ph[j].re; wa[j].im = -c * x[j] * pl->ph[j].im;
        wa[n + j].re = 0.0; wa[n + j].im = 0.0;
      }
      cfft_exec(pl->cf, wa, wb, +1);
      for (int k = 0; k < n; ++k) x[k] = wb[k].re;
    } break;
    case FFTW_RODFT01: {
      /* Y_k = Im( sum_j c_j x_j e^{+i pi (j+1)/(2n)} e^{+2 pi i (j+1)k/(2n)} ), c_{n-1} = 1 else 2;
         index the padded array by j+1 */
      for (int j = 0; j < 2 * n; ++j) { wa[j].re = 0.0; wa[j].im = 0.0; }
      for (int j = 0; j < n; ++j) {
        double c = (j == n - 1) ? 1.0 : 2.0;
        wa[j + 1].re = c * x[j] * pl->ph[j + 1].re; wa[j + 1].im = -c * x[j] * pl->ph[j + 1].im;
      }
      cfft_exec(pl->cf, wa, wb, +1);
      for (int k = 0; k < n; ++k) x[k] = wb[k].im;
    } break;
    case FFTW_REDFT11: case FFTW_RODFT11: {
      /* G_k = sum_j x_j e^{-i pi j/(2n)} e^{-2 pi i jk/(2n)};  T_k = 2 e^{-i pi (2k+1)/(4n)} G_k
         REDFT11: Y_k = Re T_k ; RODFT11: Y_k = -Im T_k */
      for (int j = 0; j < n; ++j) {
        wa[j].re = x[j] * pl->ph[j].re; wa[j].im = x[j] * pl->ph[j].im;
        wa[n + j].re = 0.0; wa[n + j].im = 0.0;
      }
      cfft_exec(pl->cf, wa, wb, -1);
      for (int k = 0; k < n; ++k) {
        cplx g = wb[k], p = pl->ph2[k];
        double tr = 2.0 * (p.re * g.re - p.im * g.im), ti = 2.0 * (p.re * g.im + p.im * g.re);
        x[k] = (pl->kind == FFTW_REDFT11) ? tr : -ti;
      }
    } break;
  }
}

/* ---- fast path: the same eight kinds through ONE complex FFT of length n/2 (n even) ----------------------------------
 * Textbook reductions, each checked against the definition path above:
 *   R2HC / HC2R        pack z_m = x_{2m} + i x_{2m+1}; X_k = E_k + w^k O_k with E, O from Z_k and conj Z_{M-k}
 *   REDFT10 / REDFT01  Makhoul: v_m = x_{2m}, v_{n-1-m} = x_{2m+1}; Y_k = 2 Re(e^{-i pi k/2n} V_k), Y_{n-k} = -2 Im(.)
 *   RODFT10 / RODFT01  through the cosine transforms: sign flip of odd inputs + reversed output, and the converse
 *   REDFT11 / RODFT11  z_m = x_{2m} + i x_{n-1-2m}, pre-twiddle e^{-i pi(4m+1)/4n}, post-twiddle e^{-i pi k/n}
 * wa, wb: scratch of at least n/2 + 1 cplx; v: scratch of n doubles. */
static void rfft_half(const r2r_plan *pl, const double *v, cplx *wa, cplx *wb) {
  /* wb[0..M] <- DFT_n(v)_k, k = 0..M (v real, length n = 2M) */
  const int M = pl->n / 2;
  for (int m = 0; m < M; ++m) { wa[m].re = v[2 * m]; wa[m].im = v[2 * m + 1]; }
  cfft_exec(pl->ch, wa, wb, -1);
  wb[M] = wb[0];
  /* in-place split needs both Z_k and Z_{M-k}: go pairwise */
  for (int k = 0; k <= M / 2; ++k) {
    const int j = M - k;
    const cplx zk = wb[k], zj = wb[j];
    /* E_k = (Z_k + conj Z_j)/2, O_k = -i (Z_k - conj Z_j)/2 ; X_k = E_k + w^k O_k ; X_j = conj(E_k) + w^j conj(O_k) */
    const double er = 0.5 * (zk.re + zj.re), ei = 0.5 * (zk.im - zj.im);
    const double orr = 0.5 * (zk.im + zj.im), oi = -0.5 * (zk.re - zj.re);
    const cplx wk = pl->wn[k], wj = pl->wn[j];
    cplx xk, xj;
    xk.re = er + (orr * wk.re - oi * wk.im); xk.im = ei + (orr * wk.im + oi * wk.re);
    xj.re = er + (orr * wj.re + oi * wj.im); xj.im = -ei + (orr * wj.im - oi * wj.re);
    wb[k] = xk; wb[j] = xj;
  }
}
static void irfft_half(const r2r_plan *pl, cplx *X, double *v, cplx *wa) {
  /* v <- n * irfft(X): X[0..M] the half spectrum (destroyed), unnormalised like HC2R */
  const int M = pl->n / 2;
  for (int k = 0; k <= M / 2; ++k) {
    const int j = M - k;
    const cplx xk = X[k], xj = X[j];
    /* Z'_k = (X_k + conj X_j) + i conj(w^k) (X_k - conj X_j) ; Z'_j likewise with k <-> j */
    const double sr = xk.re + xj.re, si = xk.im - xj.im, dr = xk.re - xj.re, di = xk.im + xj.im;
    const cplx wk = pl->wn[k], wj = pl->wn[j];
    /* conj(w) * D */
    const double ckr = dr * wk.re + di * wk.im, cki = di * wk.re - dr * wk.im;
    const double cjr = -dr * wj.re + di * wj.im, cji = di * wj.re + dr * wj.im;     /* D_j = -conj(D_k): (-dr, di) */
    cplx zk, zj;
    zk.re = sr - cki; zk.im = si + ckr;
    zj.re = sr - cji; zj.im = -si + cjr;
    if (k < M) X[k] = zk;
    if (j < M) X[j] = zj;
    if (k == 0) X[0] = zk;
  }
  cfft_exec(pl->ch, X, wa, +1);
  for (int m = 0; m < M; ++m) { v[2 * m] = wa[m].re; v[2 * m + 1] = wa[m].im; }
}

static void r2r_exec_fast(const r2r_plan *pl, double *x, cplx *wa, cplx *wb, double *v) {
  const int n = pl->n, M = n / 2;
  switch (pl->kind) {
    case FFTW_R2HC: {
      rfft_half(pl, x, wa, wb);
      for (int k = 0; k <= M; ++k) x[k] = wb[k].re;
      for (int k = 1; k < M; ++k) x[n - k] = wb[k].im;
    } break;
    case FFTW_HC2R: {
      wb[0].re = x[0]; wb[0].im = 0.0;
      for (int k = 1; k < M; ++k) { wb[k].re = x[k]; wb[k].im = x[n - k]; }
      wb[M].re = x[M]; wb[M].im = 0.0;
      irfft_half(pl, wb, x, wa);
    } break;
    case FFTW_REDFT10: case FFTW_RODFT10: {
      const int dst = (pl->kind == FFTW_RODFT10);
      for (int m = 0; m < M; ++m) {                       /* DST-II: odd inputs change sign (fft.f90:417-428) */
        v[m] = x[2 * m];
        v[n - 1 - m] = dst ? -x[2 * m + 1] : x[2 * m + 1];
      }
      rfft_half(pl, v, wa, wb);
      /* Y_k = 2 Re(q_k V_k), Y_{n-k} = -2 Im(q_k V_k), q_k = e^{-i pi k/2n}; DST-II: output reversed (fft.f90:537-560) */
      for (int k = 0; k <= M; ++k) {
        const cplx q = pl->wq[k], V = wb[k];
        const double tr = q.re * V.re - q.im * V.im, ti = q.re * V.im + q.im * V.re;
        const int a = dst ? n - 1 - k : k;
        x[a] = 2.0 * tr;
        if (k > 0 && k < M) x[dst ? k - 1 : n - k] = -2.0 * ti;
      }
    } break;
    case FFTW_REDFT01: case FFTW_RODFT01: {
      const int dst = (pl->kind == FFTW_RODFT01);
      /* DST-III: input reversed (fft.f90:859-875).  V'_k = conj(q_k) (Y_k - i Y_{n-k}), Y_n = 0 */
      for (int k = 0; k <= M; ++k) {
        const double yk = dst ? x[n - 1 - k] : x[k];
        const double yn = (k == 0) ? 0.0 : (dst ? x[k - 1] : x[n - k]);
        const cplx q = pl->wq[k];
        const double c = q.re, sg = -q.im;                  /* conj(q) = c + i sg ; (c + i sg)(yk - i yn) */
        wb[k].re = c * yk + sg * yn;
        wb[k].im = sg * yk - c * yn;
      }
      irfft_half(pl, wb, v, wa);
      for (int m = 0; m < M; ++m) {
        x[2 * m] = v[m];
        x[2 * m + 1] = dst ? -v[n - 1 - m] : v[n - 1 - m];
      }
    } break;
    case FFTW_REDFT11: case FFTW_RODFT11: {
      const int dst = (pl->kind == FFTW_RODFT11);
      for (int m = 0; m < M; ++m) {                       /* RODFT11(x)_k = (-1)^k REDFT11(reversed x)_k */
        const double a = dst ? x[n - 1 - 2 * m] : x[2 * m], b = dst ? x[2 * m] : x[n - 1 - 2 * m];
        const cplx w = pl->w4a[m];
        wa[m].re = a * w.re - b * w.im; wa[m].im = a * w.im + b * w.re;
      }
      cfft_exec(pl->ch, wa, wb, -1);
      for (int k = 0; k < M; ++k) {
        const cplx w = pl->w4b[k], V = wb[k];
        const double tr = V.re * w.re - V.im * w.im, ti = V.re * w.im + V.im * w.re;
        x[2 * k] = 2.0 * tr;
        x[n - 1 - 2 * k] = dst ? 2.0 * ti : -2.0 * ti;
      }
    } break;
  }
}

/* FFTW guru-style batched execute, in place:
 * n, stride of the transform; two howmany loops (h1n,h1s), (h2n,h2s)  (src/fft.f90:75-86,113-124) */
void oracle_r2r_execute(void *h, double *data, long stride, long h1n, long h1s, long h2n, long h2s) {
  r2r_plan *pl = (r2r_plan *)h;
  const int n = pl->n;
#pragma omp parallel
  {
    double *line = (double *)malloc(sizeof(double) * (size_t)n);
    double *vv = (double *)malloc(sizeof(double) * (size_t)(n + 2));
    cplx *wa = (cplx *)malloc(sizeof(cplx) * (size_t)(2 * n + 2));
    cplx *wb = (cplx *)malloc(sizeof(cplx) * (size_t)(2 * n + 2));
    const int fast = (pl->ch != NULL) && !g_force_definition_path;
#pragma omp for collapse(2) schedule(static)
    for (long b = 0; b < h2n; ++b)
      for (long a = 0; a < h1n; ++a) {
        double *base = data + a * h1s + b * h2s;
        if (stride == 1) {
          if (fast) r2r_exec_fast(pl, base, wa, wb, vv); else r2r_exec(pl, base, wa, wb);
        } else {
          for (int j = 0; j < n; ++j) line[j] = base[(long)j * stride];
          if (fast) r2r_exec_fast(pl, line, wa, wb, vv); else r2r_exec(pl, line, wa, wb);
          for (int j = 0; j < n; ++j) base[(long)j * stride] = line[j];
        }
      }
    free(line); free(vv); free(wa); free(wb);
  }
}

/* ------------------------------------------------------------------ initsolver pieces */

/* src/fft.f90:233-291, cell-centred ('c') table only; returns 0 on success */
int oracle_find_fft(char bc0, char bc1, char c_or_f, int *kind_fwd, int *kind_bwd, double norm[2]) {
  if (c_or_f != 'c') return 1;
  norm[0] = 2.0; norm[1] = 0.0;
  if (bc0 == 'P' && bc1 == 'P') { *kind_fwd = FFTW_R2HC; *kind_bwd = FFTW_HC2R; norm[0] = 1.0; }
  else if (bc0 == 'N' && bc1 == 'N') { *kind_fwd = FFTW_REDFT10; *kind_bwd = FFTW_REDFT01; }
  else if (bc0 == 'D' && bc1 == 'D') { *kind_fwd = FFTW_RODFT10; *kind_bwd = FFTW_RODFT01; }
  else if (bc0 == 'N' && bc1 == 'D') { *kind_fwd = FFTW_REDFT11; *kind_bwd = FFTW_REDFT11; }
  else if (bc0 == 'D' && bc1 == 'N') { *kind_fwd = FFTW_RODFT11; *kind_bwd = FFTW_RODFT11; }
  else return 2;
  return 0;
}

/* normfft as accumulated in src/fft.f90:71,87,125,150 (ix = iy = 0 for 'c') */
double oracle_normfft(int ng1, int ng2, const char bcxy[4]) {
  int kf, kb; double norm[2]; double nf = 1.0;
  oracle_find_fft(bcxy[0], bcxy[1], 'c', &kf, &kb, norm); nf = nf * norm[0] * (ng1 + norm[1] - 0);
  oracle_find_fft(bcxy[2], bcxy[3], 'c', &kf, &kb, norm); nf = nf * norm[0] * (ng2 + norm[1] - 0);
  return 1.0 / nf;
}

/* src/initsolver.f90:122-186 (CPU branch, cell-centred) ; lambda[0..n) <-> lambda(1:n) */
void oracle_eigenvalues(int n, char bc0, char bc1, double *lambda) {
  const double pi = acos(-1.0);
  for (int l = 1; l <= n; ++l) {
    double s;
    if (bc0 == 'P' && bc1 == 'P')      s = sin((1.0 * (l - 1)) * pi / (1.0 * n));
    else if (bc0 == 'N' && bc1 == 'N') s = sin((1.0 * (l - 1)) * pi / (2.0 * n));
    else if (bc0 == 'D' && bc1 == 'D') s = sin((1.0 * (l - 0)) * pi / (2.0 * n));
    else                               s = sin((1.0 * (2 * l - 1)) * pi / (4.0 * n));
    lambda[l - 1] = -4.0 * s * s;
  }
}

/* src/initsolver.f90:188-246, c_or_f = 'c'.  dzci, dzfi point at index (1-nh_d). */
void oracle_tridmatrix(char bc0, char bc1, int n, int nh_d, const double *dzci, const double *dzfi,
                       double *a, double *b, double *c) {
  const double *zc = dzci + (nh_d - 1), *zf = dzfi + (nh_d - 1);   /* zc[k] == dzci(k) */
  for (int k = 1; k <= n; ++k) {
    a[k - 1] = zf[k] * zc[k - 1];
    c[k - 1] = zf[k] * zc[k];
  }
  for (int k = 0; k < n; ++k) b[k] = -(a[k] + c[k]);
  double f0 = (bc0 == 'P') ? 0.0 : (bc0 == 'D') ? -1.0 : 1.0;
  double f1 = (bc1 == 'P') ? 0.0 : (bc1 == 'D') ? -1.0 : 1.0;
  b[0] = b[0] + f0 * a[0];
  b[n - 1] = b[n - 1] + f1 * c[n - 1];
}

/* src/initgrid.f90:17-97 with gridpoint_cluster_two_end (:102-118).
 * dzc, dzf point at index (1-nh_d), length n+2*nh_d */
void oracle_initgrid(int n, double gr, double lz, int nh_d, double *dzc_, double *dzf_) {
  double *dzc = dzc_ + (nh_d - 1), *dzf = dzf_ + (nh_d - 1);
  double *zf = (double *)malloc(sizeof(double) * (size_t)(n + 2));
  for (int k = 1; k <= n; ++k) {
    double z0 = (k - 0.0) / (1.0 * n), z;
    if (gr != 0.0) z = 0.5 * (1.0 + tanh((z0 - 0.5) * gr) / tanh(gr / 2.0)); else z = z0;
    zf[k] = z * lz;
  }
  zf[0] = 0.0;
  for (int k = 1; k <= n; ++k) dzf[k] = zf[k] - zf[k - 1];
  dzf[0] = dzf[1]; dzf[n + 1] = dzf[n];
  for (int k = 0; k <= n; ++k) dzc[k] = 0.5 * (dzf[k] + dzf[k + 1]);
  dzc[n + 1] = dzc[n];
  for (int k = 1 - nh_d; k <= 0; ++k) { dzf[k] = dzf[-k + 1]; dzc[k] = dzc[-k]; }
  for (int k = n + 1; k <= n + nh_d; ++k) { dzf[k] = dzf[2 * n - k - 1]; dzc[k] = dzc[2 * n - k]; }
  free(zf);
}

/* ------------------------------------------------------------------ Thomas, src/solver_cpu.f90:187-223 */

static void dgtsv_homebrewed(int n, const double *a, const double *b, const double *c, double *p, long ps, double *d) {
  double z = 1.0 / b[0];
  d[0] = c[0] * z;
  p[0] = p[0] * z;
  for (int l = 1; l < n - 1; ++l) {
    z = 1.0 / (b[l] - a[l] * d[l - 1]);
    d[l] = c[l] * z;
    p[l * ps] = (p[l * ps] - a[l] * p[(l - 1) * ps]) * z;
  }
  z = b[n - 1] - a[n - 1] * d[n - 2];
  if (z != 0.0) p[(n - 1) * ps] = (p[(n - 1) * ps] - a[n - 1] * p[(n - 2) * ps]) / z;
  else p[(n - 1) * ps] = 0.0;
  for (int l = n - 2; l >= 0; --l) p[l * ps] = p[l * ps] - d[l] * p[(l + 1) * ps];
}

/* The same recurrences for GB adjacent columns at a time (columns i..i+m-1 of one j: one 64-byte cache line per level
 * instead of one line per value).  Every column sees exactly the operations of dgtsv_homebrewed in the same order
 * (the library is built with -ffp-contract=off), so the results are bit-identical to the column-at-a-time sweep; only the
 * order in which memory is touched changes.  d: n*GB scratch.  lam: the m eigenvalues of the block. */
#define GB 8
static void dgtsv_block(int n, int m, const double *a, const double *b, const double *lam, const double *c, double *p, long ps,
                        double *d) {
  for (int q = 0; q < m; ++q) {
    const double z = 1.0 / (b[0] + lam[q]);
    d[q] = c[0] * z;
    p[q] = p[q] * z;
  }
  for (int l = 1; l < n - 1; ++l) {
    double *pl = p + l * ps, *pm = p + (l - 1) * ps, *dl = d + (long)l * GB, *dm = d + (long)(l - 1) * GB;
    for (int q = 0; q < m; ++q) {
      const double z = 1.0 / ((b[l] + lam[q]) - a[l] * dm[q]);
      dl[q] = c[l] * z;
      pl[q] = (pl[q] - a[l] * pm[q]) * z;
    }
  }
  {
    double *pl = p + (long)(n - 1) * ps, *pm = p + (long)(n - 2) * ps, *dm = d + (long)(n - 2) * GB;
    for (int q = 0; q < m; ++q) {
      const double z = (b[n - 1] + lam[q]) - a[n - 1] * dm[q];
      if (z != 0.0) pl[q] = (pl[q] - a[n - 1] * pm[q]) / z;
      else pl[q] = 0.0;
    }
  }
  for (int l = n - 2; l >= 0; --l) {
    double *pl = p + l * ps, *pp = p + (l + 1) * ps, *dl = d + (long)l * GB;
    for (int q = 0; q < m; ++q) pl[q] = pl[q] - dl[q] * pp[q];
  }
}

/* src/solver_cpu.f90:117-145 ; pz is (nx,ny,n) dense */
void oracle_gaussel(int nx, int ny, int n, const double *a, const double *b, const double *c,
                    const double *lambdaxy, double *pz) {
  if (!g_force_definition_path && n >= 3) {
    const int nib = (nx + GB - 1) / GB;
#pragma omp parallel
    {
      double *d = (double *)malloc(sizeof(double) * (size_t)n * GB);
#pragma omp for collapse(2) schedule(static)
      for (int j = 0; j < ny; ++j)
        for (int ib = 0; ib < nib; ++ib) {
          const int i = ib * GB, m = nx - i < GB ? nx - i : GB;
          dgtsv_block(n, m, a, b, lambdaxy + i + (long)nx * j, c, pz + i + (long)nx * j, (long)nx * ny, d);
        }
      free(d);
    }
    return;
  }
#pragma omp parallel
  {
    double *bb = (double *)malloc(sizeof(double) * (size_t)n), *d = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for collapse(2) schedule(static)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        for (int l = 0; l < n; ++l) bb[l] = b[l] + lambdaxy[i + (long)nx * j];
        dgtsv_homebrewed(n, a, bb, c, pz + i + (long)nx * j, (long)nx * ny, d);
      }
    free(bb); free(d);
  }
}

/* src/solver_cpu.f90:147-185 */
void oracle_gaussel_periodic(int nx, int ny, int n, const double *a, const double *b, const double *c,
                             const double *lambdaxy, double *pz) {
  const long ps = (long)nx * ny;
  if (!g_force_definition_path && n >= 4) {                 /* blocks of GB columns, as above: bit-identical */
    const int nib = (nx + GB - 1) / GB;
#pragma omp parallel
    {
      double *d = (double *)malloc(sizeof(double) * (size_t)n * GB);
      double *p1 = (double *)malloc(sizeof(double) * (size_t)n * GB), *p2 = (double *)malloc(sizeof(double) * (size_t)n * GB);
#pragma omp for collapse(2) schedule(static)
      for (int j = 0; j < ny; ++j)
        for (int ib = 0; ib < nib; ++ib) {
          const int i = ib * GB, m = nx - i < GB ? nx - i : GB;
          double *p = pz + i + (long)nx * j;
          const double *lam = lambdaxy + i + (long)nx * j;
          for (int l = 0; l < n - 1; ++l)
            for (int q = 0; q < m; ++q) { p1[(long)l * GB + q] = p[l * ps + q]; p2[(long)l * GB + q] = 0.0; }
          for (int q = 0; q < m; ++q) { p2[q] = -a[0]; p2[(long)(n - 2) * GB + q] = -c[n - 2]; }
          dgtsv_block(n - 1, m, a, b, lam, c, p1, GB, d);
          dgtsv_block(n - 1, m, a, b, lam, c, p2, GB, d);
          for (int q = 0; q < m; ++q)
            p[(n - 1) * ps + q] = (p[(n - 1) * ps + q] - c[n - 1] * p1[q] - a[n - 1] * p1[(long)(n - 2) * GB + q]) /
                                  ((b[n - 1] + lam[q]) + c[n - 1] * p2[q] + a[n - 1] * p2[(long)(n - 2) * GB + q]);
          for (int l = 0; l < n - 1; ++l)
            for (int q = 0; q < m; ++q) p[l * ps + q] = p1[(long)l * GB + q] + p2[(long)l * GB + q] * p[(n - 1) * ps + q];
        }
      free(d); free(p1); free(p2);
    }
    return;
  }
#pragma omp parallel
  {
    double *bb = (double *)malloc(sizeof(double) * (size_t)n), *d = (double *)malloc(sizeof(double) * (size_t)n);
    double *p1 = (double *)malloc(sizeof(double) * (size_t)n), *p2 = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for collapse(2) schedule(static)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        double *p = pz + i + (long)nx * j;
        for (int l = 0; l < n; ++l) bb[l] = b[l] + lambdaxy[i + (long)nx * j];
        for (int l = 0; l < n - 1; ++l) p1[l] = p[l * ps];
        dgtsv_homebrewed(n - 1, a, bb, c, p1, 1, d);
        for (int l = 0; l < n; ++l) p2[l] = 0.0;
        p2[0] = -a[0];
        p2[n - 2] = -c[n - 2];
        dgtsv_homebrewed(n - 1, a, bb, c, p2, 1, d);
        p[(n - 1) * ps] = (p[(n - 1) * ps] - c[n - 1] * p1[0] - a[n - 1] * p1[n - 2]) /
                          (bb[n - 1] + c[n - 1] * p2[0] + a[n - 1] * p2[n - 2]);
        for (int l = 0; l < n - 1; ++l) p[l * ps] = p1[l] + p2[l] * p[(n - 1) * ps];
      }
    free(bb); free(d); free(p1); free(p2);
  }
}

/* ------------------------------------------------------------------ solver_cpu, src/solver_cpu.f90:20-115
 * one rank: n_x = n_y = n_z = n and every transpose is the identity.
 * plans[4] = fwd-x, bwd-x, fwd-y, bwd-y (arrplan(1,1),(2,1),(1,2),(2,2)).  work: n1*n2*n3 doubles. */
void oracle_solver_cpu(const int n[3], void *const plans[4], double normfft, const double *lambdaxy,
                       const double *a, const double *b, const double *c, const char bcz[2],
                       double *p, double *work) {
  const long n1 = n[0], n2 = n[1], n3 = n[2];
  const long s1 = n1 + 2, s2 = n2 + 2;
  double *px = work;
#pragma omp parallel for collapse(2) schedule(static)
  for (long k = 0; k < n3; ++k)
    for (long j = 0; j < n2; ++j)
      memcpy(px + n1 * (j + n2 * k), p + 1 + s1 * ((j + 1) + s2 * (k + 1)), sizeof(double) * (size_t)n1);
  oracle_r2r_execute(plans[0], px, 1, n2, n1, n3, n1 * n2);          /* fwd x: fft.f90:75-86 */
  oracle_r2r_execute(plans[2], px, n1, n1, 1, n3, n1 * n2);          /* fwd y: fft.f90:113-124 */
  if (bcz[0] == 'P' && bcz[1] == 'P') oracle_gaussel_periodic((int)n1, (int)n2, (int)n3, a, b, c, lambdaxy, px);
  else                                oracle_gaussel((int)n1, (int)n2, (int)n3, a, b, c, lambdaxy, px);
  oracle_r2r_execute(plans[3], px, n1, n1, 1, n3, n1 * n2);          /* bwd y */
  oracle_r2r_execute(plans[1], px, 1, n2, n1, n3, n1 * n2);          /* bwd x */
#pragma omp parallel for collapse(2) schedule(static)
  for (long k = 0; k < n3; ++k)
    for (long j = 0; j < n2; ++j) {
      const double *src = px + n1 * (j + n2 * k);
      double *dst = p + 1 + s1 * ((j + 1) + s2 * (k + 1));
      for (long i = 0; i < n1; ++i) dst[i] = src[i] * normfft;
    }
}

/* ------------------------------------------------------------------ fillps / updt_rhs_b / correc / chkdiv */

#define UIDX(i, j, k) ((long)((i) + nh_u - 1) + su1 * ((long)((j) + nh_u - 1) + su2 * (long)((k) + nh_u - 1)))
#define PIDX(i, j, k) ((long)(i) + sp1 * ((long)(j) + sp2 * (long)(k)))

/* src/fillps.f90:42-61 */
void oracle_fillps(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                   const double *dzfi_, double dti, double rho0,
                   const double *u, const double *v, const double *w, double *p) {
  (void)dzi;
  const long su1 = nx + 2 * nh_u, su2 = ny + 2 * nh_u, sp1 = nx + 2, sp2 = ny + 2;
  const double *dzfi = dzfi_ + (nh_d - 1);
  const double dtidxi = dti * dxi, dtidyi = dti * dyi;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        double val = ((w[UIDX(i, j, k)] - w[UIDX(i, j, k - 1)]) * dti * dzfi[k] +
                      (v[UIDX(i, j, k)] - v[UIDX(i, j - 1, k)]) * dtidyi +
                      (u[UIDX(i, j, k)] - u[UIDX(i - 1, j, k)]) * dtidxi);
        p[PIDX(i, j, k)] = val * rho0;
      }
}

/* src/bound.f90:829-944, cell-centred, single rank (all six faces are domain boundaries).
 * rhsbx(ny,nz,0:1), rhsby(nx,nz,0:1), rhsbz(nx,ny,0:1) */
void oracle_updt_rhs_b(int nx, int ny, int nz, const double *rhsbx, const double *rhsby, const double *rhsbz, double *p) {
  const long sp1 = nx + 2, sp2 = ny + 2;
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j) {
      p[PIDX(1, j, k)] += rhsbx[(j - 1) + (long)ny * (k - 1)];
      p[PIDX(nx, j, k)] += rhsbx[(j - 1) + (long)ny * (k - 1) + (long)ny * nz];
    }
  for (int k = 1; k <= nz; ++k)
    for (int i = 1; i <= nx; ++i) {
      p[PIDX(i, 1, k)] += rhsby[(i - 1) + (long)nx * (k - 1)];
      p[PIDX(i, ny, k)] += rhsby[(i - 1) + (long)nx * (k - 1) + (long)nx * nz];
    }
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      p[PIDX(i, j, 1)] += rhsbz[(i - 1) + (long)nx * (j - 1)];
      p[PIDX(i, j, nz)] += rhsbz[(i - 1) + (long)nx * (j - 1) + (long)nx * ny];
    }
}

/* src/correc.f90:49-73, _CONSTANT_COEFFS_POISSON branch; rho is never dereferenced */
void oracle_correc(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                   const double *dzci_, double dt, double rho0, const double *p, double *u, double *v, double *w) {
  (void)dzi;
  const long su1 = nx + 2 * nh_u, su2 = ny + 2 * nh_u, sp1 = nx + 2, sp2 = ny + 2;
  const double *dzci = dzci_ + (nh_d - 1);
  const double rho0i = 1.0 / rho0, factori = dt * dxi, factorj = dt * dyi;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        u[UIDX(i, j, k)] = u[UIDX(i, j, k)] - factori * (p[PIDX(i + 1, j, k)] - p[PIDX(i, j, k)]) * rho0i;
        v[UIDX(i, j, k)] = v[UIDX(i, j, k)] - factorj * (p[PIDX(i, j + 1, k)] - p[PIDX(i, j, k)]) * rho0i;
        w[UIDX(i, j, k)] = w[UIDX(i, j, k)] - dt * dzci[k] * (p[PIDX(i, j, k + 1)] - p[PIDX(i, j, k)]) * rho0i;
      }
}

/* src/chkdiv.f90:46-58 (serial accumulation order k,j,i as written) */
void oracle_chkdiv(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                   const double *dzfi_, const double *u, const double *v, const double *w,
                   double *divtot, double *divmax) {
  (void)dzi;
  const long su1 = nx + 2 * nh_u, su2 = ny + 2 * nh_u;
  const double *dzfi = dzfi_ + (nh_d - 1);
  double tot = 0.0, mx = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : tot) reduction(max : mx)
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        double div = (w[UIDX(i, j, k)] - w[UIDX(i, j, k - 1)]) * dzfi[k] +
                     (v[UIDX(i, j, k)] - v[UIDX(i, j - 1, k)]) * dyi +
                     (u[UIDX(i, j, k)] - u[UIDX(i - 1, j, k)]) * dxi;
        if (fabs(div) > mx) mx = fabs(div);
        tot += div;
      }
  *divtot = tot; *divmax = mx;
}

/* src/source.f90:311-346 (single phase): the predictor's pressure-gradient term from the OLD pressure.
 *   u = u + f_t12*( - ( pold(ip)-pold(i) )*dxi )*rho0i      evaluated left to right as written */
void oracle_pres_sp_src(int nx, int ny, int nz, double f_t12, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                        const double *dzci_, double rho0i, const double *pold, double *u, double *v, double *w) {
  (void)dzi;
  const long su1 = nx + 2 * nh_u, su2 = ny + 2 * nh_u, sp1 = nx + 2, sp2 = ny + 2;
  const double *dzci = dzci_ + (nh_d - 1);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        const double pc = pold[PIDX(i, j, k)];
        u[UIDX(i, j, k)] = u[UIDX(i, j, k)] + f_t12 * (-((pold[PIDX(i + 1, j, k)] - pc) * dxi)) * rho0i;
        v[UIDX(i, j, k)] = v[UIDX(i, j, k)] + f_t12 * (-((pold[PIDX(i, j + 1, k)] - pc) * dyi)) * rho0i;
        w[UIDX(i, j, k)] = w[UIDX(i, j, k)] + f_t12 * (-((pold[PIDX(i, j, k + 1)] - pc) * dzci[k])) * rho0i;
      }
}

/* src/source.f90:247-309 (two phase), _CONSTANT_COEFFS_POISSON branch (:288-293): split pressure gradient with the
 * extrapolated pressure f1*p - f2*pold, f1 = 1 + f_t12/f_t12_o, f2 = f_t12/f_t12_o (:266-269).
 * rho(0:,0:,0:) has the same halo-1 layout as p. */
void oracle_pres_tw_src(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                        const double *dzci_, double rho0i, double f_t12, double f_t12_o, const double *p,
                        const double *pold, const double *rho, double *u, double *v, double *w) {
  (void)dzi;
  const long su1 = nx + 2 * nh_u, su2 = ny + 2 * nh_u, sp1 = nx + 2, sp2 = ny + 2;
  const double *dzci = dzci_ + (nh_d - 1);
  const double f1 = 1.0 + (f_t12 / f_t12_o), f2 = (f_t12 / f_t12_o);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        const long c = PIDX(i, j, k), cx = PIDX(i + 1, j, k), cy = PIDX(i, j + 1, k), cz = PIDX(i, j, k + 1);
        const double rhoxi = 1.0 / (0.5 * (rho[cx] + rho[c]));
        const double rhoyi = 1.0 / (0.5 * (rho[cy] + rho[c]));
        const double rhozi = 1.0 / (0.5 * (rho[cz] + rho[c]));
        const double e = f1 * p[c] - f2 * pold[c];
        u[UIDX(i, j, k)] = u[UIDX(i, j, k)] + f_t12 * ((-((p[cx] - p[c]) * dxi)) * rho0i -
                                                      (rhoxi - rho0i) * ((f1 * p[cx] - f2 * pold[cx]) - e) * dxi);
        v[UIDX(i, j, k)] = v[UIDX(i, j, k)] + f_t12 * ((-((p[cy] - p[c]) * dyi)) * rho0i -
                                                      (rhoyi - rho0i) * ((f1 * p[cy] - f2 * pold[cy]) - e) * dyi);
        w[UIDX(i, j, k)] = w[UIDX(i, j, k)] + f_t12 * ((-((p[cz] - p[c]) * dzci[k])) * rho0i -
                                                      (rhozi - rho0i) * ((f1 * p[cz] - f2 * pold[cz]) - e) * dzci[k]);
      }
}

/* pressure bookkeeping around the solve, interior only (halos untouched):
 * mode 0: pold = p          (src/apps/single_phase/main__single_phase.f90:693-699)
 * mode 1: p = pold + p      (:734-740) */
void oracle_pold_update(int nx, int ny, int nz, int mode, double *p, double *pold) {
  const long sp1 = nx + 2, sp2 = ny + 2;
  for (int k = 1; k <= nz; ++k)
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        if (mode == 0) pold[PIDX(i, j, k)] = p[PIDX(i, j, k)];
        else p[PIDX(i, j, k)] = pold[PIDX(i, j, k)] + p[PIDX(i, j, k)];
      }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
