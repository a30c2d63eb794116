// capi.cu -- implementation of include/flutas_b200.h: plan management, pointer staging and kernel
// launches for the FluTAS pressure-Poisson path on one B200 (the multi-GPU slab exchange lives in
// exchange.cuh).  Build: see flutas_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flutas_b200.h"
#include "fft_p2.h"
#include "fft_reg.h"
#include "kernels.cuh"
#include "line_plan.h"
#include "thomas_reg.cuh"
#include "thomas_uni.cuh"
#include "thomas_tile.cuh"
#include "thomas_ref.cuh"
#include "bounduvw.cuh"
#include "dz.cuh"

#include <algorithm>
#include <cmath>

using namespace fb;

#ifndef FB_PIPE_DEFAULT
#define FB_PIPE_DEFAULT 0
#endif
namespace {

thread_local std::string g_err;
std::atomic<long> g_launches{0};
cudaStream_t g_stream = nullptr;
int g_device = -1, g_rank = 0, g_nranks = 1, g_nsm = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(FLUTAS_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define LAUNCHED()                                                                                \
  do {                                                                                            \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                           \
    cudaError_t e_ = cudaGetLastError();                                                          \
    if (e_ != cudaSuccess)                                                                        \
      return fail(FLUTAS_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// optional per-stage timing with CUDA events on the library stream (bench.py's live roofline numbers)
enum Stage { ST_XF = 0, ST_YF, ST_Z, ST_YB, ST_XB, ST_FILLPS, ST_CORREC, ST_EXCH_F, ST_EXCH_B, ST_ZC, ST_ZI, ST_COUNT };
const char* const g_stage_names[ST_COUNT] = {"xfft_fwd", "yfft_fwd", "thomas_z", "yfft_bwd", "xfft_bwd",
                                             "fillps", "correc", "exchange_fwd", "exchange_bwd", "thomas_z_corr", "z_interface"};
struct StageRec { int id; cudaEvent_t e0, e1; };
bool g_prof = false;
std::vector<StageRec> g_prof_recs;
struct StageTimer {
  int id; cudaEvent_t e0 = nullptr, e1 = nullptr;
  explicit StageTimer(int id_) : id(id_) {
    if (!g_prof) return;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, g_stream);
  }
  ~StageTimer() {
    if (!e0) return;
    cudaEventRecord(e1, g_stream);
    g_prof_recs.push_back({id, e0, e1});
  }
};

int ensure_device() {
  if (g_device >= 0) return FLUTAS_B200_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(FLUTAS_B200_ERR_CUDA, "no CUDA device: libflutas_b200 has no CPU fallback (%s)",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  int cur = 0;
  CK(cudaGetDevice(&cur));
  g_device = cur;
  CK(cudaDeviceGetAttribute(&g_nsm, cudaDevAttrMultiProcessorCount, cur));
  return FLUTAS_B200_OK;
}

bool on_device(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// grow-only device scratch
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return static_cast<T*>(p); }
};

struct DevLinePlan {
  HostLinePlan h;
  DevBuf tables;
  LinePlan d{};
  int tb = 16;
  size_t smem = 0;
  // register-resident kernels (power-of-two lengths): own tables and own spectral layout
  bool use_reg = false;
  HostRegPlan hr;
  DevBuf rtables;
  RegPlan r{};
  const std::vector<int>& mode() const { return use_reg ? hr.mode : h.mode; }
  int upload_reg() {
    size_t n = hr.wN.size() + hr.wQ.size();
    for (int q = 0; q < RF_MAXPASS; ++q) n += hr.tw[q].size() + hr.tw8[q].size();
    if (int rc = rtables.reserve((n + 1) * sizeof(cpx))) return rc;
    cpx* at = rtables.as<cpx>();
    r.N = hr.N; r.M = hr.M; r.kind = hr.kind;
    for (int q = 0; q < RF_MAXPASS; ++q) {
      r.tw[q] = at; r.tw_count[q] = (int)hr.tw[q].size();
      if (!hr.tw[q].empty()) CK(cudaMemcpy(at, hr.tw[q].data(), hr.tw[q].size() * sizeof(cpx), cudaMemcpyHostToDevice));
      at += hr.tw[q].size();
    }
    for (int q = 0; q < RF_MAXPASS; ++q) {
      r.tw8[q] = hr.tw8[q].empty() ? nullptr : at;
      if (!hr.tw8[q].empty()) CK(cudaMemcpy(at, hr.tw8[q].data(), hr.tw8[q].size() * sizeof(cpx), cudaMemcpyHostToDevice));
      at += hr.tw8[q].size();
    }
    r.wN = at;
    CK(cudaMemcpy(at, hr.wN.data(), hr.wN.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    at += hr.wN.size();
    r.wQ = at;
    CK(cudaMemcpy(at, hr.wQ.data(), hr.wQ.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    return 0;
  }
  int upload() {
    const size_t nwM = h.wM.size(), nwN = h.wN.size(), nwQ = h.wQ.size(), npos = h.pos.size();
    const size_t bytes = (nwM + nwN + nwQ) * sizeof(cpx) + npos * sizeof(int);
    if (int rc = tables.reserve(bytes)) return rc;
    char* base = tables.as<char>();
    cpx* wM = (cpx*)base;
    cpx* wN = wM + nwM;
    cpx* wQ = wN + nwN;
    int* pos = (int*)(wQ + nwQ);
    CK(cudaMemcpy(wM, h.wM.data(), nwM * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(wN, h.wN.data(), nwN * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(wQ, h.wQ.data(), nwQ * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pos, h.pos.data(), npos * sizeof(int), cudaMemcpyHostToDevice));
    d.N = h.N; d.M = h.M; d.kind = h.kind; d.npass = (int)h.radix.size();
    for (int q = 0; q < d.npass; ++q) { d.radix[q] = h.radix[q]; d.sub[q] = h.sub[q]; }
    d.wM = wM; d.wN = wN; d.wQ = wQ; d.pos = pos;
    const size_t kb = 1024;
    if ((size_t)h.N * 16 * 8 <= 64 * kb) tb = 16;
    else if ((size_t)h.N * 8 * 8 <= 128 * kb) tb = 8;
    else if (fft_smem_bytes<4>(h.N) <= 220 * kb) tb = 4;
    else return fail(FLUTAS_B200_ERR_UNSUPPORTED, "transform length %d does not fit a shared-memory tile", h.N);
    smem = (tb == 16) ? fft_smem_bytes<16>(h.N) : (tb == 8) ? fft_smem_bytes<8>(h.N) : fft_smem_bytes<4>(h.N);
    return 0;
  }
};

constexpr unsigned long long PLAN_MAGIC = 0xF1074A5B200ull;

struct SolverPlan {
  unsigned long long magic = PLAN_MAGIC;
  int n1 = 0, n2 = 0;            // transform lengths (global ng1, ng2 on one rank)
  char bcxy[4] = {0, 0, 0, 0};
  DevLinePlan px, py;
  DevBuf work, scratchD, scratchP2, pstage;
  // cached coefficients
  DevBuf lam_int, abc, maps, lam_raw;
  DevBuf fft_maps;               // spectral slot -> FFTW index of both directions (flutas_b200_fft)
  int cached_nz = 0;
  bool cached_periodic = false;
  const double* key_lam = nullptr;
  std::vector<double> key_a, key_b, key_c;
  std::vector<double> key_lam_sample;
  bool cache_valid = false;
  int thomas_mode = 0;           // 0 = auto, 1 = force generic (tests)
  ThomasArgs z_uniform{};        // scalar coefficients when the z grid is exactly uniform (thomas_reg.cuh)
  unsigned long long cache_gen = 0;   // bumped whenever the cached coefficients are rebuilt
  std::vector<double> h_a, h_b, h_c;  // host copies of a, b, c (selection threshold of the reference-order columns)
  // reference-order solve of the ill-conditioned columns (thomas_ref.cuh)
  struct RefFix {
    const double* key_lam = nullptr;
    long key_ncol = 0;
    int key_nz = 0, key_singular = -1;
    bool key_periodic = false;
    unsigned long long key_gen = 0;
    double key_tol = -1.0;
    int nsel = 0;
    DevBuf col, lam, pin, z, d, piv, p2, den, F;
    RefTables T{};
  } ref, dz_ref;                 // single-rank / transposed z stage; owned columns of the distributed z solve
  unsigned long long lam_win_gen = 0;
  int lam_win_rank = -1, lam_win_n1l = 0;
  // distributed z solve (dz.cuh)
  struct DzState {
    bool ok = false;
    unsigned long long gen = ~0ull;
    int n3l = 0, rank = -1, P = 0;
    double tol = -1.0;
    ThomasArgs uni{};
    double ca = 0.0, cc = 0.0;
    int nsel = 0;
    DzOwn own{};
    DevBuf sel_dev, sel_lam_own, abc_loc;
    bool general = false;        // local blocks through the general kernel (coefficient tables) instead of the shared-LU one
    int last_used = 0;           // 1: the last solver_slab call on this plan ran the distributed z solve
  } dz;
  // slab (multi-GPU) state
  DevBuf sendrecv, pencil, lam_win;
  bool p2p = false;
  void* p2p_alloc = nullptr;                 // [pencil | recv | flags], published through CUDA IPC
  size_t p2p_bytes = 0, p2p_chunk = 0;
  double* peer_pencil[FB_MAX_RANKS] = {};
  double* peer_recv[FB_MAX_RANKS] = {};
  unsigned long long* peer_flags[FB_MAX_RANKS] = {};
  void* peer_base[FB_MAX_RANKS] = {};
  unsigned long long** d_peer_flags = nullptr;
  int* d_err = nullptr;
  int* h_err = nullptr;          // pinned mirror of d_err, refreshed behind every barrier
  unsigned long long epoch = 0;
};

struct PlanHandle {
  unsigned long long magic;
  SolverPlan* plan;
  int which;                     // 0 fwd-x, 1 bwd-x, 2 fwd-y, 3 bwd-y
  bool owned = false;            // stand-alone plan of flutas_b200_plan_r2r: the handle owns `plan`
  int dims[3] = {0, 0, 0};       // ... and knows the array it was planned for
};

SolverPlan* plan_of(void* const arrplan[4]) {
  if (!arrplan || !arrplan[0]) return nullptr;
  PlanHandle* h = (PlanHandle*)arrplan[0];
  if (h->magic != PLAN_MAGIC || !h->plan || h->plan->magic != PLAN_MAGIC) return nullptr;
  return h->plan;
}

template <int TB, bool FWD>
int launch_x(const DevLinePlan& lp, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale) {
  auto kern = xfft_kernel<TB, FWD>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
  const long nblk = (gs.nlines + TB - 1) / TB;
  kern<<<(unsigned)nblk, 256, lp.smem, g_stream>>>(lp.d, src, gs, dst, gd, scale);
  LAUNCHED();
  return 0;
}
bool g_z_uniform_ok = true;   // test hook (thomas mode 3): ignore the uniform-grid fast path
int g_fft_level = 0;   // test hook: 0 = register kernels where available, 1 = run-time-radix tile kernels only,
                       // 2 = tile kernels (power-of-two specialisations allowed) but no register kernels

bool aligned16(const void* p, const LineGeom& g) {
  return ((uintptr_t)p % 16 == 0) && (g.off0 % 2 == 0) && (g.sj % 2 == 0) && (g.sk % 2 == 0);
}

template <bool FWD>
int run_x(const DevLinePlan& lp, const double* src, LineGeom gs, double* dst, LineGeom gd, double scale) {
  if (gs.nlines >= 0x7fffffffL) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "more than 2^31 x lines on one rank");
  if (lp.use_reg) {
    if (!(FWD ? aligned16(dst, gd) : aligned16(src, gs)))
      return fail(FLUTAS_B200_ERR_ARG, "register transform kernels need a 16-byte aligned spectral work array");
    const int nsm = g_nsm > 0 ? g_nsm : 148;
    cudaError_t e = FWD ? reg_run_x_fwd(lp.r, src, gs, dst, gd, scale, nsm, g_stream)
                        : reg_run_x_bwd(lp.r, src, gs, dst, gd, scale, nsm, g_stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "xfft_reg launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
  if (g_fft_level != 1 && p2_tile_width(lp.d.N) && !kind_is_iv(lp.d.kind)) {   // (types IV: generic tile kernels only)
    cudaError_t e = p2_run_x(FWD, lp.d, src, gs, dst, gd, scale, g_stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "xfft_p2 launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
  switch (lp.tb) {
    case 16: return launch_x<16, FWD>(lp, src, gs, dst, gd, scale);
    case 8: return launch_x<8, FWD>(lp, src, gs, dst, gd, scale);
    default: return launch_x<4, FWD>(lp, src, gs, dst, gd, scale);
  }
}
template <int TB, bool FWD>
int launch_y(const DevLinePlan& lp, double* W, int n1, long n3, const SpecGeom& sg) {
  auto kern = yfft_kernel<TB, FWD>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
  const int nti = (n1 + TB - 1) / TB;
  kern<<<(unsigned)(nti * n3), 256, lp.smem, g_stream>>>(lp.d, W, n1, nti, sg);
  LAUNCHED();
  return 0;
}
template <bool FWD>
int run_y(const DevLinePlan& lp, double* W, int n1, long n3, const SpecGeom& sg) {
  if ((long)n1 * n3 >= 0x7fffffffL) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "more than 2^31 y tiles on one rank");
  if (lp.use_reg) {
    const int nsm = g_nsm > 0 ? g_nsm : 148;
    cudaError_t e = FWD ? reg_run_y_fwd(lp.r, W, n1, n3, sg, nsm, g_stream) : reg_run_y_bwd(lp.r, W, n1, n3, sg, nsm, g_stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "yfft_reg launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
  if (g_fft_level != 1 && p2_tile_width(lp.d.N) && !kind_is_iv(lp.d.kind)) {
    cudaError_t e = p2_run_y(FWD, lp.d, W, n1, n3, sg, g_stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "yfft_p2 launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
  switch (lp.tb) {
    case 16: return launch_y<16, FWD>(lp, W, n1, n3, sg);
    case 8: return launch_y<8, FWD>(lp, W, n1, n3, sg);
    default: return launch_y<4, FWD>(lp, W, n1, n3, sg);
  }
}

SpecGeom local_spec(double* W, int n1) {
  SpecGeom sg;
  for (int q = 0; q < FB_MAX_RANKS; ++q) sg.ptr[q] = W;
  sg.n1l = n1; sg.koff = 0;
  return sg;
}

int run_z(SolverPlan* sp, long ncol, int nz, const double* lam, double* W, const ColGeom* out, bool periodic, int singular);

int cache_coefficients(SolverPlan* sp, int nz, const double* lambdaxy, const double* a, const double* b, const double* c,
                       bool periodic) {
  const int n1 = sp->n1, n2 = sp->n2;
  const bool lam_dev = on_device(lambdaxy), abc_dev = on_device(a);
  bool same = sp->cache_valid && sp->cached_nz == nz && sp->key_lam == lambdaxy && sp->cached_periodic == periodic;
  if (same && !abc_dev) {
    same = (int)sp->key_a.size() == nz && !memcmp(sp->key_a.data(), a, sizeof(double) * nz) &&
           !memcmp(sp->key_b.data(), b, sizeof(double) * nz) && !memcmp(sp->key_c.data(), c, sizeof(double) * nz);
  }
  // host lambdaxy: pointer + a strided sample of 256 entries (a caller that refills the same buffer, or a recycled
  // address of a Python temporary, is caught unless it agrees on all of them; flutas_b200_solver_invalidate is the contract)
  std::vector<double> lam_sample;
  if (!lam_dev) {
    const long nl_ = (long)n1 * n2, step = nl_ > 256 ? nl_ / 256 : 1;
    for (long q = 0; q < nl_; q += step) lam_sample.push_back(lambdaxy[q]);
    lam_sample.push_back(lambdaxy[nl_ - 1]);
    if (same) same = (lam_sample.size() == sp->key_lam_sample.size()) &&
                     !memcmp(lam_sample.data(), sp->key_lam_sample.data(), lam_sample.size() * sizeof(double));
  }
  if (same) return 0;
  sp->cache_valid = false;
  sp->cache_gen += 1;
  const size_t nl = (size_t)n1 * n2;
  if (int rc = sp->lam_raw.reserve(nl * sizeof(double))) return rc;
  if (int rc = sp->lam_int.reserve(nl * sizeof(double))) return rc;
  if (int rc = sp->abc.reserve(6 * (size_t)nz * sizeof(double))) return rc;
  if (int rc = sp->maps.reserve((size_t)(n1 + n2) * sizeof(int))) return rc;
  CK(cudaMemcpyAsync(sp->lam_raw.p, lambdaxy, nl * sizeof(double), cudaMemcpyDefault, g_stream));
  double* abc = sp->abc.as<double>();
  CK(cudaMemcpyAsync(abc, a, nz * sizeof(double), cudaMemcpyDefault, g_stream));
  CK(cudaMemcpyAsync(abc + nz, b, nz * sizeof(double), cudaMemcpyDefault, g_stream));
  CK(cudaMemcpyAsync(abc + 2 * nz, c, nz * sizeof(double), cudaMemcpyDefault, g_stream));
  // second copy for the on-chip z solve: couplings that leave the system are zero unless z is periodic
  CK(cudaMemcpyAsync(abc + 3 * nz, abc, 3 * (size_t)nz * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  if (!periodic) {
    CK(cudaMemsetAsync(abc + 3 * nz, 0, sizeof(double), g_stream));                  // az[0]
    CK(cudaMemsetAsync(abc + 5 * nz + (nz - 1), 0, sizeof(double), g_stream));       // cz[nz-1]
  }
  int* maps = sp->maps.as<int>();
  CK(cudaMemcpyAsync(maps, sp->px.mode().data(), n1 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(maps + n1, sp->py.mode().data(), n2 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  permute_lambda_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, g_stream>>>(n1, n2, sp->lam_raw.as<double>(), maps,
                                                                             maps + n1, sp->lam_int.as<double>());
  LAUNCHED();
  CK(cudaStreamSynchronize(g_stream));
  {                                                      // exactly uniform z grid -> scalar coefficients for the z solve
    std::vector<double> ha(nz), hb(nz), hc(nz);
    CK(cudaMemcpy(ha.data(), a, nz * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(hb.data(), b, nz * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(hc.data(), c, nz * sizeof(double), cudaMemcpyDefault));
    thomas_detect_uniform(nz, ha.data(), hb.data(), hc.data(), periodic, sp->z_uniform);
    sp->h_a.swap(ha); sp->h_b.swap(hb); sp->h_c.swap(hc);
  }
  sp->cached_nz = nz;
  sp->cached_periodic = periodic;
  sp->key_lam = lambdaxy;
  if (!abc_dev) { sp->key_a.assign(a, a + nz); sp->key_b.assign(b, b + nz); sp->key_c.assign(c, c + nz); }
  else { sp->key_a.clear(); sp->key_b.clear(); sp->key_c.clear(); }
  sp->key_lam_sample = lam_sample;
  sp->cache_valid = true;
  return 0;
}


// z stage (solver_cpu.f90:71-77): on-chip partition/PCR kernel when nz allows it, else the generic
// scratch-field kernels (in place; a ColGeom output then needs the scatter kernel).
__global__ void scatter_cols_kernel(long ncol, int nz, const double* __restrict__ W, ColGeom og) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncol * nz) return;
  const int k = (int)(idx / ncol);
  const long col = idx - (long)k * ncol;
  const int q = k / og.n3l;
  og.ptr[q][og.koff + col + ncol * (long)(k - q * og.n3l)] = W[idx];
}

// Selection threshold of the reference-order columns: |lambda| < 4 max(|a|,|c|) * g_ref_tol (thomas_ref.cuh).
// FLUTAS_B200_REF_TOL / flutas_b200_debug_ref_tol override it; 0 switches the fix-up off.
double g_ref_tol = [] { const char* e = getenv("FLUTAS_B200_REF_TOL"); return e ? atof(e) : 1.0e-5; }();
constexpr int REF_MAX_COLUMNS = 8192;
constexpr size_t REF_MAX_SMEM = 200 * 1024;

size_t ref_smem_per_warp(int nz) { return (3 * (size_t)RefShape(nz).doubles() + 96) * sizeof(double); }

// (re)builds the tables of the selected columns for the z stage described by (lam, ncol, nz): selection on the host
// from a copy of the eigenvalues, factors on the device in the reference's operation order (once per plan / layout)
int ensure_ref(SolverPlan* sp, SolverPlan::RefFix& rf, long ncol, int nz, const double* lam, bool periodic, int singular) {
  if (rf.key_lam == lam && rf.key_ncol == ncol && rf.key_nz == nz && rf.key_periodic == periodic &&
      rf.key_singular == singular && rf.key_gen == sp->cache_gen && rf.key_tol == g_ref_tol) return 0;
  rf.key_lam = lam; rf.key_ncol = ncol; rf.key_nz = nz; rf.key_periodic = periodic; rf.key_singular = singular;
  rf.key_gen = sp->cache_gen; rf.key_tol = g_ref_tol;
  rf.nsel = 0;
  if (!(g_ref_tol > 0.0) || (int)sp->h_a.size() != nz || ref_smem_per_warp(nz) > REF_MAX_SMEM || ncol > 0x7fffffffL) return 0;
  double amax = 0.0;
  for (int k = 0; k < nz; ++k) amax = std::max(amax, std::max(std::fabs(sp->h_a[k]), std::fabs(sp->h_c[k])));
  const double thr = 4.0 * amax * g_ref_tol;
  std::vector<double> hl((size_t)ncol);
  CK(cudaMemcpyAsync(hl.data(), lam, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  std::vector<int> sel;
  for (long q = 0; q < ncol; ++q)
    if (std::fabs(hl[q]) < thr) sel.push_back((int)q);
  if ((int)sel.size() > REF_MAX_COLUMNS) {                 // keep the most ill-conditioned ones
    std::nth_element(sel.begin(), sel.begin() + REF_MAX_COLUMNS, sel.end(),
                     [&](int x, int y) { return std::fabs(hl[x]) < std::fabs(hl[y]); });
    sel.resize(REF_MAX_COLUMNS);
    std::sort(sel.begin(), sel.end());
  }
  const int ns = (int)sel.size();
  if (ns == 0) return 0;
  std::vector<double> sl(ns);
  std::vector<unsigned char> pin(ns);
  for (int q = 0; q < ns; ++q) { sl[q] = hl[sel[q]]; pin[q] = (singular && sl[q] == 0.0) ? 1 : 0; }
  const size_t tab = (size_t)ns * nz * sizeof(double);
  if (int rc = rf.col.reserve(ns * sizeof(int))) return rc;
  if (int rc = rf.lam.reserve(ns * sizeof(double))) return rc;
  if (int rc = rf.pin.reserve(ns)) return rc;
  if (int rc = rf.z.reserve(tab)) return rc;
  if (int rc = rf.d.reserve(tab)) return rc;
  if (int rc = rf.F.reserve(tab)) return rc;
  if (int rc = rf.piv.reserve(ns * sizeof(double))) return rc;
  if (periodic) {
    if (int rc = rf.p2.reserve(tab)) return rc;
    if (int rc = rf.den.reserve(ns * sizeof(double))) return rc;
  }
  CK(cudaMemcpyAsync(rf.col.p, sel.data(), ns * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(rf.lam.p, sl.data(), ns * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(rf.pin.p, pin.data(), ns, cudaMemcpyHostToDevice, g_stream));
  RefTables& T = rf.T;
  T.nsel = ns; T.nz = nz; T.m = periodic ? nz - 1 : nz; T.periodic = periodic ? 1 : 0;
  const double* abc = sp->abc.as<double>();               // a | b | c exactly as passed by the caller
  T.a = abc; T.b = abc + nz; T.c = abc + 2 * nz;
  T.col = rf.col.as<int>(); T.lam = rf.lam.as<double>(); T.pin = rf.pin.as<unsigned char>();
  T.z = rf.z.as<double>(); T.d = rf.d.as<double>(); T.piv = rf.piv.as<double>();
  T.p2 = periodic ? rf.p2.as<double>() : nullptr; T.den = periodic ? rf.den.as<double>() : nullptr;
  ref_factor_kernel<<<(unsigned)((ns + 63) / 64), 64, 0, g_stream>>>(T);
  LAUNCHED();
  CK(cudaStreamSynchronize(g_stream));                     // sel / sl / pin leave scope
  rf.nsel = ns;
  return 0;
}

int run_z_main(SolverPlan* sp, long ncol, int nz, const double* lam, double* W, const ColGeom* out, bool periodic, int singular);

// z stage = [reference-order solve of the ill-conditioned columns into a side buffer] + main kernel (all columns, in
// place or scattered to the peers) + [scatter of the side buffer over the main kernel's result]
int run_z(SolverPlan* sp, long ncol, int nz, const double* lam, double* W, const ColGeom* out, bool periodic, int singular) {
  if (int rc = ensure_ref(sp, sp->ref, ncol, nz, lam, periodic, singular)) return rc;
  SolverPlan::RefFix& rf = sp->ref;
  if (rf.nsel) {
    const size_t per = ref_smem_per_warp(nz);
    int warps = (int)(REF_MAX_SMEM / per);
    warps = warps > 4 ? 4 : warps;
    static size_t configured = 0;
    if (per * warps > configured) {
      CK(cudaFuncSetAttribute(ref_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per * warps)));
      configured = per * warps;
    }
    ref_solve_kernel<<<(unsigned)((rf.nsel + warps - 1) / warps), 32 * warps, per * warps, g_stream>>>(rf.T, ncol, W, rf.F.as<double>());
    LAUNCHED();
  }
  if (int rc = run_z_main(sp, ncol, nz, lam, W, out, periodic, singular)) return rc;
  if (rf.nsel) {
    ColGeom og;
    if (out) og = *out;
    else { for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = W; og.n3l = nz; og.koff = 0; }
    const long cnt = (long)rf.nsel * nz;
    ref_scatter_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(rf.T, ncol, rf.F.as<double>(), og);
    LAUNCHED();
  }
  return 0;
}

int run_z_main(SolverPlan* sp, long ncol, int nz, const double* lam, double* W, const ColGeom* out, bool periodic, int singular) {
  const double* abc = sp->abc.as<double>();
  bool done = false;
  static const bool uni_env = [] { const char* e = getenv("FLUTAS_B200_THOMAS_UNI"); return !(e && e[0] == '0'); }();
  // exactly uniform z grid: shared LU factors + TMA tile loads (thomas_uni.cuh).  B200, z stage: 1024^3 6.38 -> 4.56 ms,
  // 1024x512x512 0.97 -> 0.86 ms, 1024x1024x512 1.80 -> 1.74 ms against the general kernel's best tile shape.
  if (sp->thomas_mode == 0 && g_z_uniform_ok && uni_env && sp->z_uniform.uniform) {
    int rc = thomas_uni_run(ncol, nz, lam, W, W, out, periodic, singular, g_nsm > 0 ? g_nsm : 148, &sp->z_uniform, g_stream, &done);
    if (rc) return fail(FLUTAS_B200_ERR_CUDA, "thomas_uni launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (!done && sp->thomas_mode == 0) {                      // register-resident kernel, tiles through TMA (L = 16 shapes)
    int rc = thomas_reg_tma_run(ncol, nz, abc + 3 * nz, abc + 4 * nz, abc + 5 * nz, lam, W, W, out, periodic, singular,
                                g_nsm > 0 ? g_nsm : 148, g_z_uniform_ok ? &sp->z_uniform : nullptr, g_stream, &done);
    if (rc) return fail(FLUTAS_B200_ERR_CUDA, "thomas_reg_tma launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (!done && sp->thomas_mode == 0) {                      // register-resident persistent kernel (cp.async slots)
    int rc = thomas_reg_run(ncol, nz, abc + 3 * nz, abc + 4 * nz, abc + 5 * nz, lam, W, W, out, periodic, singular,
                            g_nsm > 0 ? g_nsm : 148, g_z_uniform_ok ? &sp->z_uniform : nullptr, g_stream, &done);
    if (rc) return fail(FLUTAS_B200_ERR_CUDA, "thomas_reg launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (!done && (sp->thomas_mode == 0 || sp->thomas_mode == 2)) {   // shared-memory tile kernel
    int rc = thomas_tile_run(ncol, nz, abc + 3 * nz, abc + 4 * nz, abc + 5 * nz, lam, W, out, periodic, singular, g_stream, &done);
    if (rc) return fail(FLUTAS_B200_ERR_CUDA, "thomas_tile launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (done) g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (done) return 0;
  const size_t npts = (size_t)ncol * nz;
  if (int rc = sp->scratchD.reserve(npts * sizeof(double))) return rc;
  const unsigned nb = (unsigned)((ncol + 127) / 128);
  if (periodic) {
    if (int rc = sp->scratchP2.reserve(npts * sizeof(double))) return rc;
    thomas_periodic_generic_kernel<<<nb, 128, 0, g_stream>>>(ncol, nz, abc, abc + nz, abc + 2 * nz, lam, W,
                                                             sp->scratchD.as<double>(), sp->scratchP2.as<double>(), singular);
  } else {
    thomas_generic_kernel<<<nb, 128, 0, g_stream>>>(ncol, nz, abc, abc + nz, abc + 2 * nz, lam, W,
                                                    sp->scratchD.as<double>(), singular);
  }
  LAUNCHED();
  if (out) {
    scatter_cols_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, g_stream>>>(ncol, nz, W, *out);
    LAUNCHED();
  }
  return 0;
}

// ---- multi-GPU slab exchange ------------------------------------------------------------------
// schedule knobs of flutas_b200_solver_slab (flutas_b200_slab_config; defaults from the environment)
int g_pipe_chunks = [] { const char* e = getenv("FLUTAS_B200_PIPE"); return e ? atoi(e) : FB_PIPE_DEFAULT; }();
int g_pipe_xsm = [] { const char* e = getenv("FLUTAS_B200_PIPE_XSM"); return e ? atoi(e) : 50; }();
int g_zcopy = [] { const char* e = getenv("FLUTAS_B200_ZCOPY"); return e ? atoi(e) : -1; }();
int g_dz = [] { const char* e = getenv("FLUTAS_B200_DZ"); return e ? atoi(e) : 1; }();   // distributed z solve where the shape allows it
flutas_b200_alltoall_fn g_a2a = nullptr;
void* g_a2a_ctx = nullptr;
flutas_b200_halo_fn g_halo = nullptr;
void* g_halo_ctx = nullptr;

struct P2PBlob {                       // what one rank publishes to the others
  cudaIpcMemHandle_t handle;
  unsigned long long bytes;
  unsigned long long chunk_doubles;
};

// cross-GPU barrier through flag words in peer memory: rank r stores `epoch` into slot r of every
// peer's flag array, then waits until all slots of its own array have reached `epoch`.
__global__ void p2p_barrier_kernel(int rank, int nranks, unsigned long long epoch, unsigned long long* own,
                                   unsigned long long* const* peers, int* err, long long limit) {
  const int q = threadIdx.x;
  if (q >= nranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers[q] + rank), "l"(epoch) : "memory");
  const long long t0 = clock64();
  unsigned long long v = 0;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(own + q) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > limit) { atomicAdd(err, 1); break; }          // never hang the GPU: the host sees `err`
  } while (true);
}

StencilGeom stencil_geom(int nx, int ny, int nz, int nh_u) {
  StencilGeom g;
  g.nx = nx; g.ny = ny; g.nz = nz; g.nh_u = nh_u;
  g.su1 = nx + 2 * nh_u; g.su2 = ny + 2 * nh_u; g.sp1 = nx + 2; g.sp2 = ny + 2;
  return g;
}

// staging of host fields for the stencil entry points
DevBuf g_stage[6], g_coef, g_red;

struct FieldRef {            // a field that is either already on the device or staged into g_stage[slot]
  double* dev = nullptr;
  const double* host = nullptr;
  size_t bytes = 0;
  bool staged = false;
};
int stage_in(FieldRef& f, int slot, const double* ptr, size_t count, bool copy_in) {
  f.bytes = count * sizeof(double);
  if (on_device(ptr)) { f.dev = const_cast<double*>(ptr); f.staged = false; return 0; }
  if (int rc = g_stage[slot].reserve(f.bytes)) return rc;
  f.dev = g_stage[slot].as<double>(); f.host = ptr; f.staged = true;
  if (copy_in) CK(cudaMemcpyAsync(f.dev, ptr, f.bytes, cudaMemcpyHostToDevice, g_stream));
  return 0;
}
int stage_out(FieldRef& f) {
  if (f.staged) CK(cudaMemcpyAsync(const_cast<double*>(f.host), f.dev, f.bytes, cudaMemcpyDeviceToHost, g_stream));
  return 0;
}

}  // namespace

extern "C" {

const char* flutas_b200_version(void) { return "flutas_b200 0.1 (sm_100a, FP64 pressure-Poisson path)"; }
const char* flutas_b200_last_error(void) { return g_err.c_str(); }
long flutas_b200_launch_count(void) { return g_launches.load(); }

int flutas_b200_profile_enable(int on) {
  g_prof = (on != 0);
  return FLUTAS_B200_OK;
}

int flutas_b200_profile_stage_count(void) { return ST_COUNT; }
const char* flutas_b200_profile_stage_name(int id) { return (id >= 0 && id < ST_COUNT) ? g_stage_names[id] : ""; }

// Synchronises the stream, then returns for every stage the summed device time (ms) and the number of
// timed launches since the last read; clears the records.
int flutas_b200_profile_read(double* ms_sum, long* counts) {
  if (int rc = ensure_device()) return rc;
  CK(cudaStreamSynchronize(g_stream));
  for (int q = 0; q < ST_COUNT; ++q) { ms_sum[q] = 0.0; counts[q] = 0; }
  for (StageRec& r : g_prof_recs) {
    float ms = 0.f;
    CK(cudaEventSynchronize(r.e1));
    CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_sum[r.id] += ms; counts[r.id] += 1;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof_recs.clear();
  return FLUTAS_B200_OK;
}

int flutas_b200_init(int device, int rank, int nranks) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(FLUTAS_B200_ERR_CUDA, "no CUDA device: libflutas_b200 has no CPU fallback (%s)",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(FLUTAS_B200_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(FLUTAS_B200_ERR_ARG, "bad rank %d of %d", rank, nranks);
  CK(cudaSetDevice(device));
  g_device = device; g_rank = rank; g_nranks = nranks;
  CK(cudaDeviceGetAttribute(&g_nsm, cudaDevAttrMultiProcessorCount, device));
  return FLUTAS_B200_OK;
}

int flutas_b200_set_stream(void* s) { g_stream = (cudaStream_t)s; return FLUTAS_B200_OK; }

void* flutas_b200_alloc(size_t bytes) {
  if (ensure_device()) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { fail(FLUTAS_B200_ERR_CUDA, "cudaMalloc(%zu) failed", bytes); return nullptr; }
  return p;
}
// managed memory: valid on the host (un-ported Fortran routines keep working on the same arrays, pages migrate) and on
// the device (every entry point of this library then runs in place, asynchronously) -- what the reference's GPU build uses
// for all its fields (main__single_phase.f90:157-163)
void* flutas_b200_alloc_managed(size_t bytes) {
  if (ensure_device()) return nullptr;
  void* p = nullptr;
  if (cudaMallocManaged(&p, bytes, cudaMemAttachGlobal) != cudaSuccess) { fail(FLUTAS_B200_ERR_CUDA, "cudaMallocManaged(%zu) failed", bytes); return nullptr; }
  return p;
}
void flutas_b200_free(void* p) { if (p) cudaFree(p); }
int flutas_b200_memcpy(void* dst, const void* src, size_t bytes) {
  if (int rc = ensure_device()) return rc;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}
int flutas_b200_synchronize(void) {
  if (int rc = ensure_device()) return rc;
  CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

}  // extern "C"
// tables of one transform direction (shared-memory tile kernels always, register kernels where the length allows)
static int build_line(DevLinePlan& lp, int N, int kind) {
  lp.h = make_line_plan(N, kind);
  if (!lp.h.ok) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "transform length %d: need an even length whose half factors into 2,3,5", N);
  if (int rc = lp.upload()) return rc;
  if (g_fft_level != 0 || !reg_fft_supported(lp.h.N)) return 0;
  lp.hr = make_reg_plan(lp.h.N, lp.h.kind);
  lp.use_reg = lp.hr.ok;
  return lp.use_reg ? lp.upload_reg() : 0;
}
extern "C" {
int flutas_b200_fftini(const int n_x[3], const int n_y[3], const char bcxy[4], const char c_or_f[2],
                       void* arrplan[4], double* normfft) {
  if (!n_x || !n_y || !bcxy || !c_or_f || !arrplan || !normfft) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  for (int q = 0; q < 4; ++q) arrplan[q] = nullptr;
  if (c_or_f[0] != 'c' || c_or_f[1] != 'c')
    return fail(FLUTAS_B200_ERR_UNSUPPORTED, "only cell-centred ('c') transforms are on this path (FluTAS passes 'c','c','c')");
  const int kx = kind_from_bc(bcxy[0], bcxy[1]), ky = kind_from_bc(bcxy[2], bcxy[3]);
  if (kx < 0 || ky < 0)
    return fail(FLUTAS_B200_ERR_UNSUPPORTED, "pressure BC pair %c%c/%c%c: the transform table of src/fft.f90:233-291 has "
                "PP, NN, DD, ND, DN only (sanity.f90:212-222)", bcxy[0], bcxy[1], bcxy[2], bcxy[3]);
  if (int rc = ensure_device()) return rc;
  SolverPlan* sp = new SolverPlan();
  sp->n1 = n_x[0]; sp->n2 = n_y[1];
  memcpy(sp->bcxy, bcxy, 4);
  int rc = build_line(sp->px, sp->n1, kx);
  if (!rc) rc = build_line(sp->py, sp->n2, ky);
  if (rc) { delete sp; return rc; }
  // normfft exactly as src/fft.f90:71,87,125,150 (norm = (1,0) for PP, (2,0) for NN/DD; ix = iy = 0)
  double nf = 1.0;
  nf = nf * (kx == KIND_PP ? 1.0 : 2.0) * (sp->n1 + 0.0 - 0);
  nf = nf * (ky == KIND_PP ? 1.0 : 2.0) * (sp->n2 + 0.0 - 0);
  *normfft = 1.0 / nf;
  for (int q = 0; q < 4; ++q) { PlanHandle* h = new PlanHandle(); h->magic = PLAN_MAGIC; h->plan = sp; h->which = q; arrplan[q] = h; }
  return FLUTAS_B200_OK;
}

static void destroy_solver_plan(SolverPlan* sp) {
  for (DevBuf* b : {&sp->px.rtables, &sp->py.rtables}) b->release();
  for (DevBuf* b : {&sp->px.tables, &sp->py.tables, &sp->work, &sp->scratchD, &sp->scratchP2, &sp->pstage,
                    &sp->lam_int, &sp->abc, &sp->maps, &sp->lam_raw, &sp->fft_maps}) b->release();
  for (DevBuf* b : {&sp->sendrecv, &sp->pencil, &sp->lam_win}) b->release();
  for (DevBuf* b : {&sp->dz.sel_dev, &sp->dz.sel_lam_own, &sp->dz.abc_loc}) b->release();
  for (DevBuf* b : {&sp->dz_ref.col, &sp->dz_ref.lam, &sp->dz_ref.pin, &sp->dz_ref.z, &sp->dz_ref.d, &sp->dz_ref.piv, &sp->dz_ref.p2, &sp->dz_ref.den, &sp->dz_ref.F}) b->release();
  for (DevBuf* b : {&sp->ref.col, &sp->ref.lam, &sp->ref.pin, &sp->ref.z, &sp->ref.d, &sp->ref.piv, &sp->ref.p2, &sp->ref.den, &sp->ref.F}) b->release();
  for (int q = 0; q < FB_MAX_RANKS; ++q)
    if (sp->peer_base[q] && sp->peer_base[q] != sp->p2p_alloc) cudaIpcCloseMemHandle(sp->peer_base[q]);
  if (sp->p2p_alloc) cudaFree(sp->p2p_alloc);
  if (sp->d_peer_flags) cudaFree(sp->d_peer_flags);
  if (sp->d_err) cudaFree(sp->d_err);
  if (sp->h_err) cudaFreeHost(sp->h_err);
  sp->magic = 0;
  delete sp;
}

int flutas_b200_fftend(void* arrplan[4]) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp) return fail(FLUTAS_B200_ERR_ARG, "fftend: not a flutas_b200 plan");
  if (((PlanHandle*)arrplan[0])->owned) return fail(FLUTAS_B200_ERR_ARG, "fftend: stand-alone plans end with flutas_b200_destroy_plan");
  destroy_solver_plan(sp);
  for (int q = 0; q < 4; ++q) { delete (PlanHandle*)arrplan[q]; arrplan[q] = nullptr; }
  return FLUTAS_B200_OK;
}

int flutas_b200_solver_invalidate(void* const arrplan[4]) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp) return fail(FLUTAS_B200_ERR_ARG, "not a flutas_b200 plan");
  sp->cache_valid = false;
  return FLUTAS_B200_OK;
}

// test hook, read by fftini (register kernels) and by every solve (tile kernels): 0 = register-resident kernels
// for power-of-two lengths, 1 = always the run-time-radix tile kernels, 2 = tile kernels incl. their
// power-of-two specialisations but no register kernels
int flutas_b200_debug_generic_fft(int level) {
  g_fft_level = level;
  return FLUTAS_B200_OK;
}

// test hook: 0 = automatic choice of the z solver (register kernel, then the shared-memory tile kernel, then
// the generic ones), 1 = always the generic (scratch-field) kernels, 2 = skip the register kernel
int flutas_b200_debug_thomas_mode(void* const arrplan[4], int mode) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp) return fail(FLUTAS_B200_ERR_ARG, "not a flutas_b200 plan");
  g_z_uniform_ok = (mode != 3);                          // 3 = register kernel with coefficient tables even on a uniform grid
  sp->thomas_mode = (mode == 3) ? 0 : mode;
  return FLUTAS_B200_OK;
}

// test hook: selection tolerance of the reference-order columns (0 = off; a huge value selects every column up to the cap)
int flutas_b200_debug_ref_tol(double tol) {
  g_ref_tol = tol;
  return FLUTAS_B200_OK;
}

int flutas_b200_solver(const int n[3], void* const arrplan[4], double normfft, const double* lambdaxy,
                       const double* a, const double* b, const double* c, const char bcz[2],
                       const char c_or_f[3], double* p) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp) return fail(FLUTAS_B200_ERR_ARG, "solver: arrplan was not created by flutas_b200_fftini");
  if (!n || !lambdaxy || !a || !b || !c || !bcz || !c_or_f || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (c_or_f[2] != 'c') return fail(FLUTAS_B200_ERR_UNSUPPORTED, "c_or_f(3) must be 'c'");
  if (g_nranks != 1) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "multi-rank solve: use flutas_b200_solver_slab");
  const int n1 = n[0], n2 = n[1], n3 = n[2];
  if (n1 != sp->n1 || n2 != sp->n2) return fail(FLUTAS_B200_ERR_ARG, "n = (%d,%d,%d) does not match the plan (%d,%d)", n1, n2, n3, sp->n1, sp->n2);
  const bool periodic = (bcz[0] == 'P' && bcz[1] == 'P');
  if (n3 < 2 || (periodic && n3 < 4)) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "n3 = %d too small", n3);
  const bool zsing = periodic || (bcz[0] == 'N' && bcz[1] == 'N');
  const bool xysing = (sp->bcxy[0] != 'D' && sp->bcxy[2] != 'D');   // PP or NN in both x and y
  const int singular = (zsing && xysing) ? 1 : 0;

  if (int rc = cache_coefficients(sp, n3, lambdaxy, a, b, c, periodic)) return rc;
  const size_t npts = (size_t)n1 * n2 * n3;
  if (int rc = sp->work.reserve(npts * sizeof(double))) return rc;
  double* W = sp->work.as<double>();

  const size_t pcount = (size_t)(n1 + 2) * (n2 + 2) * (n3 + 2);
  double* pd = p;
  const bool host_p = !on_device(p);
  if (host_p) {
    if (int rc = sp->pstage.reserve(pcount * sizeof(double))) return rc;
    pd = sp->pstage.as<double>();
    CK(cudaMemcpyAsync(pd, p, pcount * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  }

  const long s1 = n1 + 2, s2 = n2 + 2;
  LineGeom gp{1 + s1 * (1 + s2), s1, s1 * s2, n2, (long)n2 * n3};
  LineGeom gw{0, (long)n1, (long)n1 * n2, n2, (long)n2 * n3};
  const double* abc = sp->abc.as<double>();
  const double* lam = sp->lam_int.as<double>();

  { StageTimer t(ST_XF); if (int rc = run_x<true>(sp->px, pd, gp, W, gw, 1.0)) return rc; }   // solver_cpu.f90:59
  const SpecGeom sg = local_spec(W, n1);
  { StageTimer t(ST_YF); if (int rc = run_y<true>(sp->py, W, n1, n3, sg)) return rc; }         // :65
  { StageTimer t(ST_Z); if (int rc = run_z(sp, (long)n1 * n2, n3, lam, W, nullptr, periodic, singular)) return rc; }   // :71-77
  { StageTimer t(ST_YB); if (int rc = run_y<false>(sp->py, W, n1, n3, sg)) return rc; }        // :86
  { StageTimer t(ST_XB); if (int rc = run_x<false>(sp->px, W, gw, pd, gp, normfft)) return rc; }  // :89,93

  if (host_p) {
    CK(cudaMemcpyAsync(p, pd, pcount * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  return FLUTAS_B200_OK;
}

// fft(plan,arr), src/fft.f90:181-193: one batched, unnormalised, in-place r2r transform of a dense pencil array in FFTW's
// element order.  The solver never calls this (its stages hand the spectrum over in slot order); it exists for hosts that
// keep the reference's solver_cpu.f90 and for the per-kind parity tests.  n = extents of arr.
int flutas_b200_fft(void* plan, const int n[3], double* arr) {
  PlanHandle* h = (PlanHandle*)plan;
  if (!h || h->magic != PLAN_MAGIC || !h->plan || h->plan->magic != PLAN_MAGIC) return fail(FLUTAS_B200_ERR_ARG, "fft: not a flutas_b200 plan");
  if (!n || !arr) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  SolverPlan* sp = h->plan;
  const bool xdir = h->which < 2, fwd = (h->which % 2) == 0;
  DevLinePlan& lp = xdir ? sp->px : sp->py;
  const int n1 = n[0], n2 = n[1];
  const long n3 = n[2];
  if (n1 < 1 || n2 < 1 || n3 < 1 || lp.h.N != (xdir ? n1 : n2))
    return fail(FLUTAS_B200_ERR_ARG, "fft: array (%d,%d,%ld) does not match the plan (length %d along %c)", n1, n2, n3, lp.h.N, xdir ? 'x' : 'y');
  const size_t npts = (size_t)n1 * n2 * n3;
  if (int rc = sp->work.reserve(npts * sizeof(double))) return rc;
  double* W = sp->work.as<double>();
  double* A = arr;
  const bool host = !on_device(arr);
  if (host) {
    if (int rc = sp->pstage.reserve(npts * sizeof(double))) return rc;
    A = sp->pstage.as<double>();
    CK(cudaMemcpyAsync(A, arr, npts * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  }
  const size_t nmx = sp->px.mode().size(), nmy = sp->py.mode().size();
  if (!sp->fft_maps.p) {
    if (int rc = sp->fft_maps.reserve((nmx + nmy + 1) * sizeof(int))) return rc;
    if (nmx) CK(cudaMemcpyAsync(sp->fft_maps.p, sp->px.mode().data(), nmx * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    if (nmy) CK(cudaMemcpyAsync(sp->fft_maps.as<int>() + nmx, sp->py.mode().data(), nmy * sizeof(int), cudaMemcpyHostToDevice, g_stream));
  }
  const int* map = sp->fft_maps.as<int>() + (xdir ? 0 : nmx);
  const unsigned nblk = (unsigned)((npts + 255) / 256);
  const LineGeom gw{0, (long)n1, (long)n1 * n2, n2, (long)n2 * n3};
  int rc = 0;
  if (xdir) {
    if (fwd) {
      rc = run_x<true>(lp, A, gw, W, gw, 1.0);
      if (!rc) { fft_permute_kernel<<<nblk, 256, 0, g_stream>>>(n1, n2, n3, 0, 1, map, W, A); LAUNCHED(); }
    } else {
      fft_permute_kernel<<<nblk, 256, 0, g_stream>>>(n1, n2, n3, 0, 0, map, A, W); LAUNCHED();
      rc = run_x<false>(lp, W, gw, A, gw, 1.0);
    }
  } else {
    const SpecGeom sg = local_spec(W, n1);
    if (fwd) {
      CK(cudaMemcpyAsync(W, A, npts * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
      rc = run_y<true>(lp, W, n1, n3, sg);
      if (!rc) { fft_permute_kernel<<<nblk, 256, 0, g_stream>>>(n1, n2, n3, 1, 1, map, W, A); LAUNCHED(); }
    } else {
      fft_permute_kernel<<<nblk, 256, 0, g_stream>>>(n1, n2, n3, 1, 0, map, A, W); LAUNCHED();
      rc = run_y<false>(lp, W, n1, n3, sg);
      if (!rc) CK(cudaMemcpyAsync(A, W, npts * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    }
  }
  if (rc) return rc;
  if (host) {
    CK(cudaMemcpyAsync(arr, A, npts * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  return FLUTAS_B200_OK;
}

// Stand-alone plan with the arguments of fftw_plan_guru_r2r as src/fft.f90:75-86,113-124 passes them: rank 1,
// (n, is) of the transform, two howmany dimensions (n, is), in place (os = is).  Served: the x layout (is = 1, dense lines)
// and the y layout (is = howmany(1).n, howmany(1).is = 1).  kind = FFTW's integer code (src/fftw.f90:41-61).
int flutas_b200_plan_r2r(int n, int is, const int hm_n[2], const int hm_is[2], int kind, void** plan) {
  if (!hm_n || !hm_is || !plan) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  *plan = nullptr;
  int k = -1, bwd = 0;
  switch (kind) {
    case 0: k = KIND_PP; break;             // R2HC
    case 1: k = KIND_PP; bwd = 1; break;    // HC2R
    case 5: k = KIND_NN; break;             // REDFT10
    case 4: k = KIND_NN; bwd = 1; break;    // REDFT01
    case 9: k = KIND_DD; break;             // RODFT10
    case 8: k = KIND_DD; bwd = 1; break;    // RODFT01
    case 6: k = KIND_ND; break;             // REDFT11 (its own inverse up to 2n)
    case 10: k = KIND_DN; break;            // RODFT11
    default: return fail(FLUTAS_B200_ERR_UNSUPPORTED, "r2r kind %d: the cell-centred table of src/fft.f90:233-291 uses R2HC, HC2R, "
                         "REDFT10/01, RODFT10/01, REDFT11, RODFT11", kind);
  }
  int xdir = -1, dims[3] = {0, 0, hm_n[1]};
  if (is == 1 && hm_is[0] == n && (long)hm_is[1] == (long)n * hm_n[0]) { xdir = 1; dims[0] = n; dims[1] = hm_n[0]; }
  else if (hm_is[0] == 1 && is == hm_n[0] && (long)hm_is[1] == (long)hm_n[0] * n) { xdir = 0; dims[0] = hm_n[0]; dims[1] = n; }
  if (xdir < 0 || n < 2 || hm_n[0] < 1 || hm_n[1] < 1)
    return fail(FLUTAS_B200_ERR_UNSUPPORTED, "plan_r2r: only the dense x (is = 1) and y (is = n1) pencil layouts of src/fft.f90:75-86,113-124");
  if (int rc = ensure_device()) return rc;
  SolverPlan* sp = new SolverPlan();
  if (xdir) sp->n1 = n; else sp->n2 = n;
  if (int rc = build_line(xdir ? sp->px : sp->py, n, k)) { delete sp; return rc; }
  PlanHandle* h = new PlanHandle();
  h->magic = PLAN_MAGIC; h->plan = sp; h->which = (xdir ? 0 : 2) + bwd; h->owned = true;
  for (int q = 0; q < 3; ++q) h->dims[q] = dims[q];
  *plan = h;
  return FLUTAS_B200_OK;
}

int flutas_b200_plan_dims(void* plan, int n[3]) {
  PlanHandle* h = (PlanHandle*)plan;
  if (!h || h->magic != PLAN_MAGIC || !h->owned || !n) return fail(FLUTAS_B200_ERR_ARG, "plan_dims: not a stand-alone flutas_b200 plan");
  for (int q = 0; q < 3; ++q) n[q] = h->dims[q];
  return FLUTAS_B200_OK;
}

int flutas_b200_destroy_plan(void* plan) {
  PlanHandle* h = (PlanHandle*)plan;
  if (!h || h->magic != PLAN_MAGIC || !h->owned || !h->plan) return fail(FLUTAS_B200_ERR_ARG, "destroy_plan: not a stand-alone flutas_b200 plan");
  destroy_solver_plan(h->plan);
  h->magic = 0;
  delete h;
  return FLUTAS_B200_OK;
}

int flutas_b200_set_alltoall(flutas_b200_alltoall_fn fn, void* ctx) {
  g_a2a = fn;
  g_a2a_ctx = ctx;
  return FLUTAS_B200_OK;
}

// Schedule of the slab solver: pipe_chunks = number of k-chunks of the pipelined forward half (x transform of chunk c+1
// under the y transform + NVLink stores of chunk c; 0/1 = off), pipe_xsm_pct = share of the SMs given to the x kernels,
// zcopy = 1 / 0 forces the copy-engine / fused backward exchange (-1 = the built-in rule).  A negative value leaves a
// knob unchanged.  All ranks must use the same values (the host layer tunes them collectively: SlabComm.autotune).
int flutas_b200_slab_config(int pipe_chunks, int pipe_xsm_pct, int zcopy) {
  if (pipe_chunks >= 0) g_pipe_chunks = pipe_chunks;
  if (pipe_xsm_pct >= 0) g_pipe_xsm = pipe_xsm_pct;
  if (zcopy >= -1 && zcopy <= 1) g_zcopy = zcopy;
  return FLUTAS_B200_OK;
}

// which z stage the last flutas_b200_solver_slab call on this plan used: 1 = distributed (dz.cuh), 0 = transposes
int flutas_b200_slab_last_distributed(void* const arrplan[4]) {
  SolverPlan* sp = plan_of(arrplan);
  return sp ? sp->dz.last_used : 0;
}

// 1 (default): on a slab decomposition with direct NVLink exchange and an exactly uniform z grid the z stage runs as a
// distributed tridiagonal solve (dz.cuh) and the two all-to-all transposes disappear; 0: always transpose.  Same value on
// every rank.
int flutas_b200_slab_distributed_z(int on) {
  g_dz = on ? 1 : 0;
  return FLUTAS_B200_OK;
}

size_t flutas_b200_p2p_handle_bytes(void) { return sizeof(P2PBlob); }

// Allocates this rank's exchange memory [pencil | recv | flags] and returns its IPC handle in `blob`.
int flutas_b200_p2p_export(void* const arrplan[4], const int n_local[3], void* blob) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp || !n_local || !blob) return fail(FLUTAS_B200_ERR_ARG, "p2p_export: bad arguments");
  const int P = g_nranks;
  if (P < 2 || P > FB_MAX_RANKS) return fail(FLUTAS_B200_ERR_ARG, "p2p needs 2..%d ranks (flutas_b200_init)", FB_MAX_RANKS);
  if (sp->n1 % P) return fail(FLUTAS_B200_ERR_ARG, "ng1 = %d is not divisible by %d ranks", sp->n1, P);
  const size_t chunk = (size_t)(sp->n1 / P) * sp->n2 * n_local[2];
  const size_t field = chunk * P * sizeof(double);
  const size_t bytes = 2 * field + 4096;
  if (sp->p2p_alloc) { cudaFree(sp->p2p_alloc); sp->p2p_alloc = nullptr; }
  CK(cudaMalloc(&sp->p2p_alloc, bytes));
  CK(cudaMemset(sp->p2p_alloc, 0, bytes));
  sp->p2p_bytes = bytes; sp->p2p_chunk = chunk;
  P2PBlob b;
  memset(&b, 0, sizeof(b));
  CK(cudaIpcGetMemHandle(&b.handle, sp->p2p_alloc));
  b.bytes = bytes; b.chunk_doubles = chunk;
  memcpy(blob, &b, sizeof(b));
  return FLUTAS_B200_OK;
}

// `blobs`: the nranks blobs in rank order (all-gathered by the host).  Maps every peer's exchange memory.
int flutas_b200_p2p_attach(void* const arrplan[4], const void* blobs) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp || !blobs || !sp->p2p_alloc) return fail(FLUTAS_B200_ERR_ARG, "p2p_attach: call p2p_export first");
  const int P = g_nranks;
  const P2PBlob* bl = (const P2PBlob*)blobs;
  const size_t field = sp->p2p_chunk * P * sizeof(double);
  for (int q = 0; q < P; ++q) {
    if (bl[q].chunk_doubles != sp->p2p_chunk) return fail(FLUTAS_B200_ERR_ARG, "rank %d has a different chunk size", q);
    void* base = sp->p2p_alloc;
    if (q != g_rank) {
      cudaIpcMemHandle_t h = bl[q].handle;
      cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(FLUTAS_B200_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
    }
    sp->peer_base[q] = base;
    sp->peer_pencil[q] = (double*)base;
    sp->peer_recv[q] = (double*)((char*)base + field);
    sp->peer_flags[q] = (unsigned long long*)((char*)base + 2 * field);
  }
  if (!sp->d_peer_flags) CK(cudaMalloc(&sp->d_peer_flags, FB_MAX_RANKS * sizeof(void*)));
  CK(cudaMemcpy(sp->d_peer_flags, sp->peer_flags, FB_MAX_RANKS * sizeof(void*), cudaMemcpyHostToDevice));
  if (!sp->d_err) { CK(cudaMalloc(&sp->d_err, sizeof(int))); CK(cudaMemset(sp->d_err, 0, sizeof(int))); }
  if (!sp->h_err) { CK(cudaHostAlloc((void**)&sp->h_err, sizeof(int), cudaHostAllocDefault)); *sp->h_err = 0; }
  sp->epoch = 0;
  sp->p2p = true;
  return FLUTAS_B200_OK;
}

// number of barrier time-outs seen so far (0 = healthy); synchronises the stream
int flutas_b200_p2p_errors(void* const arrplan[4]) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp || !sp->d_err) return 0;
  int v = 0;
  cudaStreamSynchronize(g_stream);
  cudaMemcpy(&v, sp->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  return v;
}


}  // extern "C"

namespace {

// A peer that has not arrived after FLUTAS_B200_P2P_TIMEOUT_S seconds (default 10; SM clock taken as 2 GHz) makes the
// barrier give up and count an error.  The count is mirrored into pinned host memory right behind every barrier, so
// the NEXT solver_slab call on the plan (and flutas_b200_p2p_errors at any time) fails instead of returning stale data.
static int p2p_barrier(SolverPlan* sp) {
  static const long long limit = [] {
    const char* e = getenv("FLUTAS_B200_P2P_TIMEOUT_S");
    const double sec = e ? atof(e) : 10.0;
    return (long long)((sec > 0.01 ? sec : 0.01) * 2.0e9);
  }();
  sp->epoch += 1;
  p2p_barrier_kernel<<<1, 32, 0, g_stream>>>(g_rank, g_nranks, sp->epoch, sp->peer_flags[g_rank], sp->d_peer_flags, sp->d_err, limit);
  LAUNCHED();
  if (sp->h_err) CK(cudaMemcpyAsync(sp->h_err, sp->d_err, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  return 0;
}

// true when the distributed z solve can serve this plan / decomposition (the same answer on every rank)
bool dz_shape_ok(const SolverPlan* sp, int n1, int n2, int n3l, int P) {
  const long ncol = (long)n1 * n2;
  const bool uni = g_z_uniform_ok && sp->z_uniform.uniform;          // shared-LU kernel; else the general kernel (coefficient tables)
  if (!uni && !thomas_reg_local_ok(n3l, ncol)) return false;
  return g_dz && sp->p2p && sp->thomas_mode == 0 && P >= 2 && P <= FB_DZ_MAXG &&
         n3l % 16 == 0 && n3l / 16 >= 2 && n3l / 16 <= 32 && (ncol % 2) == 0 && (ncol % P) == 0 && ((ncol / P) % 2) == 0;
}

struct DzLayout {                 // offsets (doubles) inside every rank's [pencil | recv] exchange allocation
  long ncol, ncol_own;
  size_t oYF, oYL, oPF, oPL, oQF, oQL, oXP, oXN;     // pencil region
  size_t oGATH, oOVR;                                 // recv region
};
DzLayout dz_layout(long ncol, int P, int nsel, int ng3) {
  DzLayout L;
  L.ncol = ncol; L.ncol_own = ncol / P;
  L.oYF = 0; L.oYL = (size_t)ncol; L.oPF = 2 * (size_t)ncol; L.oPL = 3 * (size_t)ncol; L.oQF = 4 * (size_t)ncol;
  L.oQL = 5 * (size_t)ncol; L.oXP = 6 * (size_t)ncol; L.oXN = 7 * (size_t)ncol;
  L.oGATH = 0; L.oOVR = (size_t)nsel * ng3 + ((size_t)nsel * ng3 & 1);
  return L;
}
DzPeers dz_peers(double* const* base, int P, size_t off) {
  DzPeers d;
  for (int q = 0; q < FB_DZ_MAXG; ++q) d.p[q] = (q < P) ? base[q] + off : nullptr;
  return d;
}

// one sweep of the rank-local block over the slab W (ncol x n3l): shared-LU kernel on an exactly uniform grid, else the
// general kernel with the local coefficient tables; corr = nullptr: pass 1, else pass 2
int dz_local_run(SolverPlan* sp, long ncol, int n3l, double* W, int sing_loc, const ThomasCorr* corr, bool* done) {
  SolverPlan::DzState& dz = sp->dz;
  const int nsm = g_nsm > 0 ? g_nsm : 148;
  const double* lam = sp->lam_int.as<double>();
  int rc;
  if (dz.general) {
    const double* t = dz.abc_loc.as<double>();
    rc = thomas_reg_local_run(ncol, n3l, t, t + n3l, t + 2 * n3l, lam, W, sing_loc, nsm, corr, g_stream, done);
  } else {
    rc = thomas_uni_local_run(ncol, n3l, lam, W, sing_loc, nsm, &dz.uni, corr, g_stream, done);
  }
  if (rc) return fail(FLUTAS_B200_ERR_CUDA, "distributed z: local solve failed: %s", cudaGetErrorString((cudaError_t)rc));
  if (*done) g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// once per plan / coefficient generation: local block description, the reference-order column list, and the
// right-hand-side independent interface coefficients p, q (two local solves on unit right-hand sides)
int dz_setup(SolverPlan* sp, int n1, int n2, int n3l, int P, int r, bool periodic, int singular) {
  SolverPlan::DzState& dz = sp->dz;
  const bool general = !(g_z_uniform_ok && sp->z_uniform.uniform);   // the same choice on every rank (global property)
  if (dz.gen == sp->cache_gen && dz.n3l == n3l && dz.rank == r && dz.P == P && dz.tol == g_ref_tol && dz.general == general) return 0;
  dz.gen = sp->cache_gen; dz.n3l = n3l; dz.rank = r; dz.P = P; dz.tol = g_ref_tol;
  dz.ok = false;
  const int ng3 = n3l * P, k0 = r * n3l;
  const long ncol = (long)n1 * n2;
  const size_t nloc = (size_t)ncol * n3l;
  if ((int)sp->h_a.size() != ng3) return 0;
  dz.general = general;
  if (!dz.general) {
    if (!thomas_detect_uniform(n3l, sp->h_a.data() + k0, sp->h_b.data() + k0, sp->h_c.data() + k0, false, dz.uni)) return 0;
  } else {                                                 // local rows of a, b, c with the couplings out of the block cut
    std::vector<double> t(3 * (size_t)n3l);
    for (int k = 0; k < n3l; ++k) { t[k] = sp->h_a[k0 + k]; t[n3l + k] = sp->h_b[k0 + k]; t[2 * n3l + k] = sp->h_c[k0 + k]; }
    t[0] = 0.0; t[3 * (size_t)n3l - 1] = 0.0;
    if (int rc = dz.abc_loc.reserve(t.size() * sizeof(double))) return rc;
    CK(cudaMemcpyAsync(dz.abc_loc.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  dz.ca = (periodic || r > 0) ? sp->h_a[k0] : 0.0;
  dz.cc = (periodic || r < P - 1) ? sp->h_c[k0 + n3l - 1] : 0.0;
  const double* lam = sp->lam_int.as<double>();
  // reference-order columns: the same list on every rank (the eigenvalues are identical everywhere)
  std::vector<int> sel;
  std::vector<double> hl((size_t)ncol);
  CK(cudaMemcpyAsync(hl.data(), lam, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  if (g_ref_tol > 0.0 && ref_smem_per_warp(ng3) <= REF_MAX_SMEM) {
    double amax = 0.0;
    for (int k = 0; k < ng3; ++k) amax = std::max(amax, std::max(std::fabs(sp->h_a[k]), std::fabs(sp->h_c[k])));
    const double thr = 4.0 * amax * g_ref_tol;
    for (long q = 0; q < ncol; ++q) if (std::fabs(hl[q]) < thr) sel.push_back((int)q);
    if ((int)sel.size() > REF_MAX_COLUMNS) {
      std::nth_element(sel.begin(), sel.begin() + REF_MAX_COLUMNS, sel.end(),
                       [&](int x, int y) { return std::fabs(hl[x]) < std::fabs(hl[y]); });
      sel.resize(REF_MAX_COLUMNS);
      std::sort(sel.begin(), sel.end());
    }
  }
  dz.nsel = (int)sel.size();
  if (8 * (size_t)ncol > nloc || (size_t)dz.nsel * (size_t)(ng3 + n3l) + 2 > nloc) return 0;     // exchange areas do not fit
  const long ncol_own = ncol / P;
  for (int q = 0; q <= P; ++q) dz.own.own0[q] = (int)(std::lower_bound(sel.begin(), sel.end(), (int)std::min<long>(q * ncol_own, ncol)) - sel.begin());
  for (int q = P + 1; q <= FB_DZ_MAXG; ++q) dz.own.own0[q] = dz.nsel;
  const int nown = dz.own.own0[r + 1] - dz.own.own0[r];
  if (dz.nsel) {
    if (int rc = dz.sel_dev.reserve(dz.nsel * sizeof(int))) return rc;
    CK(cudaMemcpyAsync(dz.sel_dev.p, sel.data(), dz.nsel * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    std::vector<double> lo((size_t)std::max(nown, 1));
    for (int q = 0; q < nown; ++q) lo[q] = hl[sel[dz.own.own0[r] + q]];
    if (int rc = dz.sel_lam_own.reserve(lo.size() * sizeof(double))) return rc;
    CK(cudaMemcpyAsync(dz.sel_lam_own.p, lo.data(), lo.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (nown) {                                            // factor tables of the owned columns, global coefficients
      if (int rc = ensure_ref(sp, sp->dz_ref, nown, ng3, dz.sel_lam_own.as<double>(), periodic, singular)) return rc;
      if (sp->dz_ref.nsel != nown) return fail(FLUTAS_B200_ERR_ARG, "distributed z: reference-order selection mismatch (%d != %d)", sp->dz_ref.nsel, nown);
    }
  }
  // p = T_g^{-1}(a_first e_first), q = T_g^{-1}(c_last e_last) at the first / last level -> the column owners
  const DzLayout L = dz_layout(ncol, P, dz.nsel, ng3);
  double* S = sp->peer_recv[r];                            // scratch: the recv region is free outside a solve
  const int nsm = g_nsm > 0 ? g_nsm : 148;
  const int sing_loc = (singular && r == P - 1) ? 1 : 0;
  const unsigned nb = (unsigned)((ncol + 255) / 256);
  for (int pass = 0; pass < 2; ++pass) {
    CK(cudaMemsetAsync(S, 0, nloc * sizeof(double), g_stream));
    dz_fill_plane_kernel<<<nb, 256, 0, g_stream>>>(pass == 0 ? S : S + (size_t)ncol * (n3l - 1), ncol, pass == 0 ? dz.ca : dz.cc);
    LAUNCHED();
    bool done = false;
    if (int rc = dz_local_run(sp, ncol, n3l, S, sing_loc, nullptr, &done)) return rc;
    if (!done) return 0;
    dz_send_planes_kernel<<<nb, 256, 0, g_stream>>>(ncol, ncol_own, r, S, S + (size_t)ncol * (n3l - 1),
                                                     dz_peers(sp->peer_pencil, P, pass == 0 ? L.oPF : L.oQF),
                                                     dz_peers(sp->peer_pencil, P, pass == 0 ? L.oPL : L.oQL));
    LAUNCHED();
  }
  // every rank's p / q planes have landed and nobody's scratch (= the area peers gather into) is in use any more
  if (int rc = p2p_barrier(sp)) return rc;
  dz.ok = true;
  return 0;
}

// the z stage of the slab solver without transposes; W1 = this rank's slab (n1, n2, n3l) after the forward y transform
int dz_solve_z(SolverPlan* sp, int n1, int n2, int n3l, int P, int r, double* W1, bool periodic, int singular) {
  SolverPlan::DzState& dz = sp->dz;
  const int ng3 = n3l * P, k0 = r * n3l;
  const long ncol = (long)n1 * n2, ncol_own = ncol / P;
  const DzLayout L = dz_layout(ncol, P, dz.nsel, ng3);
  const double* lam = sp->lam_int.as<double>();
  const int nsm = g_nsm > 0 ? g_nsm : 148;
  const int sing_loc = (singular && r == P - 1) ? 1 : 0;
  const unsigned nb = (unsigned)((ncol + 255) / 256);
  double* mine = sp->peer_pencil[r];
  double* mine_recv = sp->peer_recv[r];
  const int nown = dz.own.own0[r + 1] - dz.own.own0[r];
  bool done = false;
  {
    StageTimer t(ST_Z);
    if (dz.nsel) {                                         // the reference-order columns, whole, to their owners (before y overwrites b)
      const long cnt = (long)dz.nsel * n3l;
      dz_gather_sel_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(dz.nsel, n3l, k0, ncol, ncol_own, dz.sel_dev.as<int>(), W1,
                                                                                dz.own, dz_peers(sp->peer_recv, P, L.oGATH));
      LAUNCHED();
    }
    if (int rc = dz_local_run(sp, ncol, n3l, W1, sing_loc, nullptr, &done)) return rc;                        // pass 1
    if (!done) return fail(FLUTAS_B200_ERR_CUDA, "distributed z: local solve not served for this shape");
  }
  {
    StageTimer t(ST_ZI);
    dz_send_planes_kernel<<<nb, 256, 0, g_stream>>>(ncol, ncol_own, r, W1, W1 + (size_t)ncol * (n3l - 1),
                                                     dz_peers(sp->peer_pencil, P, L.oYF), dz_peers(sp->peer_pencil, P, L.oYL));
    LAUNCHED();
    if (int rc = p2p_barrier(sp)) return rc;
    dz_interface_kernel<<<(unsigned)((ncol_own + 127) / 128), 128, 0, g_stream>>>(P, periodic ? 1 : 0, ncol_own, (long)r * ncol_own, mine + L.oPF, mine + L.oPL,
                                                                                   mine + L.oQF, mine + L.oQL, mine + L.oYF, mine + L.oYL,
                                                                                   dz_peers(sp->peer_pencil, P, L.oXP), dz_peers(sp->peer_pencil, P, L.oXN));
    LAUNCHED();
    if (nown) {
      SolverPlan::RefFix& rf = sp->dz_ref;
      const size_t per = ref_smem_per_warp(ng3);
      int warps = (int)(REF_MAX_SMEM / per);
      warps = warps > 4 ? 4 : warps;
      static size_t configured = 0;
      if (per * warps > configured) {
        CK(cudaFuncSetAttribute(ref_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per * warps)));
        configured = per * warps;
      }
      ref_solve_kernel<<<(unsigned)((nown + warps - 1) / warps), 32 * warps, per * warps, g_stream>>>(rf.T, (long)nown, mine_recv + L.oGATH, rf.F.as<double>());
      LAUNCHED();
      const long cnt = (long)nown * ng3;
      dz_push_sel_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(nown, dz.own.own0[r], ng3, n3l, rf.F.as<double>(),
                                                                              dz_peers(sp->peer_recv, P, L.oOVR));
      LAUNCHED();
    }
    if (int rc = p2p_barrier(sp)) return rc;
  }
  {
    StageTimer t(ST_ZC);
    const ThomasCorr corr{mine + L.oXP, mine + L.oXN, dz.ca, dz.cc};
    if (int rc = dz_local_run(sp, ncol, n3l, W1, sing_loc, &corr, &done)) return rc;                           // pass 2
    if (!done) return fail(FLUTAS_B200_ERR_CUDA, "distributed z: correction solve not served for this shape");
    if (dz.nsel) {
      const long cnt = (long)dz.nsel * n3l;
      dz_apply_sel_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(dz.nsel, n3l, ncol, dz.sel_dev.as<int>(), mine_recv + L.oOVR, W1);
      LAUNCHED();
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int flutas_b200_solver_slab(const int n[3], void* const arrplan[4], double normfft, const double* lambdaxy_global,
                            const double* a, const double* b, const double* c, const char bcz[2],
                            const char c_or_f[3], double* p) {
  SolverPlan* sp = plan_of(arrplan);
  if (!sp) return fail(FLUTAS_B200_ERR_ARG, "solver_slab: arrplan was not created by flutas_b200_fftini");
  if (!n || !lambdaxy_global || !a || !b || !c || !bcz || !c_or_f || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (c_or_f[2] != 'c') return fail(FLUTAS_B200_ERR_UNSUPPORTED, "c_or_f(3) must be 'c'");
  const int P = g_nranks, r = g_rank;
  if (P == 1) return flutas_b200_solver(n, arrplan, normfft, lambdaxy_global, a, b, c, bcz, c_or_f, p);
  if (P > FB_MAX_RANKS) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "at most %d ranks (one NVSwitch box)", FB_MAX_RANKS);
  const int n1 = n[0], n2 = n[1], n3l = n[2], ng3 = n3l * P;
  if (n1 != sp->n1 || n2 != sp->n2) return fail(FLUTAS_B200_ERR_ARG, "n = (%d,%d,%d) does not match the plan (%d,%d)", n1, n2, n3l, sp->n1, sp->n2);
  if (n1 % P) return fail(FLUTAS_B200_ERR_ARG, "ng1 = %d is not divisible by %d ranks (sanity.f90:159-167)", n1, P);
  if (!sp->p2p && !g_a2a) return fail(FLUTAS_B200_ERR_ARG, "multi-rank solve needs flutas_b200_set_alltoall or flutas_b200_p2p_attach");
  const int n1l = n1 / P;
  const bool periodic = (bcz[0] == 'P' && bcz[1] == 'P');
  if (ng3 < 4) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "ng3 = %d too small", ng3);
  const bool zsing = periodic || (bcz[0] == 'N' && bcz[1] == 'N');
  const bool xysing = (sp->bcxy[0] != 'D' && sp->bcxy[2] != 'D');
  const int singular = (zsing && xysing) ? 1 : 0;

  if (sp->p2p && sp->h_err && *sp->h_err)
    return fail(FLUTAS_B200_ERR_CUDA, "solver_slab: %d cross-GPU barrier time-out(s) in an earlier solve on this plan -- a peer did not "
                "arrive within the limit (FLUTAS_B200_P2P_TIMEOUT_S); the pressure of that solve is invalid", *sp->h_err);
  if (int rc = cache_coefficients(sp, ng3, lambdaxy_global, a, b, c, periodic)) return rc;
  const size_t chunk = (size_t)n1l * n2 * n3l, nloc = chunk * P;
  if (sp->lam_win_gen != sp->cache_gen || sp->lam_win_rank != r || sp->lam_win_n1l != n1l ||
      sp->lam_win.cap < (size_t)n1l * n2 * sizeof(double)) {
    sp->lam_win_gen = sp->cache_gen; sp->lam_win_rank = r; sp->lam_win_n1l = n1l;
    if (int rc = sp->lam_win.reserve((size_t)n1l * n2 * sizeof(double))) return rc;
    // this rank's x rows of the permuted eigenvalues: lam_win(i_l, ry) = lam_int(r*n1l + i_l, ry); constant until the
    // coefficient cache is invalidated
    CK(cudaMemcpy2DAsync(sp->lam_win.p, (size_t)n1l * sizeof(double), sp->lam_int.as<double>() + (size_t)r * n1l,
                         (size_t)n1 * sizeof(double), (size_t)n1l * sizeof(double), n2, cudaMemcpyDeviceToDevice, g_stream));
  }
  if (int rc = sp->work.reserve(nloc * sizeof(double))) return rc;
  double* W1 = sp->work.as<double>();
  double *S = nullptr, *W2 = nullptr, *R = nullptr;
  if (sp->p2p) {
    if (chunk != sp->p2p_chunk) return fail(FLUTAS_B200_ERR_ARG, "local size changed since p2p_export");
    W2 = sp->peer_pencil[r]; R = sp->peer_recv[r];
  } else {
    if (int rc = sp->sendrecv.reserve(nloc * sizeof(double))) return rc;
    if (int rc = sp->pencil.reserve(nloc * sizeof(double))) return rc;
    S = sp->sendrecv.as<double>(); R = S; W2 = sp->pencil.as<double>();
  }

  const size_t pcount = (size_t)(n1 + 2) * (n2 + 2) * (n3l + 2);
  double* pd = p;
  const bool host_p = !on_device(p);
  if (host_p) {
    if (int rc = sp->pstage.reserve(pcount * sizeof(double))) return rc;
    pd = sp->pstage.as<double>();
    CK(cudaMemcpyAsync(pd, p, pcount * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  }
  const long s1 = n1 + 2, s2 = n2 + 2;
  LineGeom gp{1 + s1 * (1 + s2), s1, s1 * s2, n2, (long)n2 * n3l};
  LineGeom gw{0, (long)n1, (long)n1 * n2, n2, (long)n2 * n3l};

  // Distributed z solve (dz.cuh): every transform stays on the rank's own slab, the z stage is two local sweeps around a
  // 2P x 2P interface system per column, and 32 bytes per COLUMN cross NVLink instead of 14 bytes per POINT.
  if (dz_shape_ok(sp, n1, n2, n3l, P) && chunk == sp->p2p_chunk) {
    if (int rc = dz_setup(sp, n1, n2, n3l, P, r, periodic, singular)) return rc;
    sp->dz.last_used = sp->dz.ok ? 1 : 0;
    if (sp->dz.ok) {
      { StageTimer t(ST_XF); if (int rc = run_x<true>(sp->px, pd, gp, W1, gw, 1.0)) return rc; }
      const SpecGeom sg = local_spec(W1, n1);
      { StageTimer t(ST_YF); if (int rc = run_y<true>(sp->py, W1, n1, n3l, sg)) return rc; }
      if (int rc = dz_solve_z(sp, n1, n2, n3l, P, r, W1, periodic, singular)) return rc;
      { StageTimer t(ST_YB); if (int rc = run_y<false>(sp->py, W1, n1, n3l, sg)) return rc; }
      { StageTimer t(ST_XB); if (int rc = run_x<false>(sp->px, W1, gw, pd, gp, normfft)) return rc; }
      if (host_p) {
        CK(cudaMemcpyAsync(p, pd, pcount * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (sp->h_err && *sp->h_err)
          return fail(FLUTAS_B200_ERR_CUDA, "solver_slab: cross-GPU barrier time-out (%d): a peer did not arrive; p is invalid", *sp->h_err);
      }
      return FLUTAS_B200_OK;
    }
  }

  sp->dz.last_used = 0;
  // Forward half, pipelined over k-chunks (direct-store exchange only): the y transform + NVLink stores of chunk c run on
  // part of the SMs while the x transform of chunk c+1 runs on the rest (second stream), so the HBM-bound x stage hides
  // behind the NVLink-bound y stage.  FLUTAS_B200_PIPE = number of chunks (0/1 = off), FLUTAS_B200_PIPE_XSM = percentage
  // of the SMs given to the x kernels.  Measured at 2 GPUs (profiles/r01_pipe_forward_N2.log): 512^3 1.560 -> 1.521 ms,
  // 1024^3 12.46 -> 12.57..12.84 ms: the y kernel with remote stores needs all SMs itself (its remote stores are not a pure
  // link-bound tail), so splitting the SMs only divides the throughput -> off by default.
  const int pipe_chunks = g_pipe_chunks;
  const int pipe_xsm = g_pipe_xsm < 10 ? 10 : g_pipe_xsm > 90 ? 90 : g_pipe_xsm;
  const bool pipelined = sp->p2p && pipe_chunks > 1 && pipe_chunks <= 16 && (n3l % pipe_chunks) == 0 && sp->px.use_reg && sp->py.use_reg;
  if (pipelined) {
    static cudaStream_t s_x = nullptr;
    static cudaEvent_t ev_in = nullptr, ev_c[16];
    if (!s_x) {
      CK(cudaStreamCreateWithFlags(&s_x, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
      for (int c = 0; c < 16; ++c) CK(cudaEventCreateWithFlags(&ev_c[c], cudaEventDisableTiming));
    }
    struct Scoped {                                      // launch configuration of run_x / run_y for this scope
      cudaStream_t s0; int n0;
      Scoped(cudaStream_t s, int nsm) : s0(g_stream), n0(g_nsm) { g_stream = s; g_nsm = nsm; }
      ~Scoped() { g_stream = s0; g_nsm = n0; }
    };
    const int nsm_all = g_nsm > 0 ? g_nsm : 148;
    int nsm_x = nsm_all * pipe_xsm / 100; if (nsm_x < 1) nsm_x = 1;
    const int nsm_y = nsm_all - nsm_x > 0 ? nsm_all - nsm_x : 1;
    const int n3c = n3l / pipe_chunks;
    SpecGeom sg;
    for (int q = 0; q < FB_MAX_RANKS; ++q) sg.ptr[q] = nullptr;
    for (int q = 0; q < P; ++q) sg.ptr[q] = sp->peer_pencil[q];
    sg.n1l = n1l;
    StageTimer t(ST_YF);                                 // the whole pipelined forward half is booked on the y stage
    CK(cudaEventRecord(ev_in, g_stream));
    CK(cudaStreamWaitEvent(s_x, ev_in, 0));
    y_wide_request() = 1;
    int rc = 0;
    for (int c = 0; c < pipe_chunks && !rc; ++c) {
      LineGeom gpc = gp, gwc = gw;
      gpc.off0 += (long)c * n3c * gp.sk; gwc.off0 += (long)c * n3c * gw.sk;
      gpc.nlines = gwc.nlines = (long)n2 * n3c;
      { Scoped cfg(s_x, nsm_x); rc = run_x<true>(sp->px, pd, gpc, W1, gwc, 1.0); }
      if (rc) break;
      CK(cudaEventRecord(ev_c[c], s_x));
      CK(cudaStreamWaitEvent(g_stream, ev_c[c], 0));
      sg.koff = (long)r * (long)chunk + (long)c * n3c * (long)n1l * n2;
      { Scoped cfg(g_stream, nsm_y); rc = run_y<true>(sp->py, W1 + (size_t)c * n3c * n1 * n2, n1, n3c, sg); }
    }
    y_wide_request() = 0;
    if (rc) return rc;
  } else {
  { StageTimer t(ST_XF); if (int rc = run_x<true>(sp->px, pd, gp, W1, gw, 1.0)) return rc; }
  {
    // y transform whose store IS the pack (NCCL) or the exchange itself (direct stores into peer pencils)
    SpecGeom sg;
    for (int q = 0; q < FB_MAX_RANKS; ++q) sg.ptr[q] = nullptr;
    for (int q = 0; q < P; ++q) sg.ptr[q] = sp->p2p ? sp->peer_pencil[q] : S + (size_t)q * chunk;
    sg.n1l = n1l; sg.koff = sp->p2p ? (long)r * (long)chunk : 0;
    StageTimer t(ST_YF);
    y_wide_request() = sp->p2p ? 1 : 0;                  // remote stores: prefer 128-byte row pieces
    const int rc = run_y<true>(sp->py, W1, n1, n3l, sg);
    y_wide_request() = 0;
    if (rc) return rc;
  }
  }
  {
    StageTimer t(ST_EXCH_F);
    if (sp->p2p) { if (int rc = p2p_barrier(sp)) return rc; }
    else if (g_a2a(g_a2a_ctx, S, W2, chunk * sizeof(double), (void*)g_stream)) return fail(FLUTAS_B200_ERR_CUDA, "all-to-all callback failed");
  }
  // Backward exchange of the direct path, two ways.  (a) fused: the z kernel scatters every level of its tile straight into
  // the owners' receive buffers (128-byte pieces, one per level, `level stride` = 8 ncol bytes apart).  (b) copy: the z
  // kernel solves in place and the slab of levels owned by peer q -- ONE contiguous block of the pencil, and contiguous in
  // q's receive buffer too -- goes by cudaMemcpyAsync (copy engines) on its own stream.
  // Measured (profiles/r01_scaling_v10_N8.jsonl, r01_zcopy_N2.log): (a) overlaps transfer and solve and wins at 1024^3
  // (8 GPUs: z + send 1.44 ms = 650 GB/s; 2 GPUs: 3.34 ms against 6.28 ms for (b)); at 2048 x 2048 x 1024 on 8 GPUs it
  // collapses to 188 GB/s (20 ms + 5 ms of barrier skew) while the same grid on 2 GPUs still moves 400 GB/s.  What sets
  // that case apart is the number of distinct remote 2 MB pages a tile's stores touch at once: 896 (7 peers x 128 levels,
  // 4 MB level stride) against 448-512 in every case that runs well -- consistent with a remote-translation working set
  // of ~512 entries.  Hence (b) when a tile touches more than 600 remote pages.  That rule rests on ONE data point and (b)
  // has been verified for correctness and timed at 2 GPUs only (the round's GPU budget ended there);
  // FLUTAS_B200_ZCOPY = 0 / 1 forces either mode.
  const int zcopy_env = g_zcopy;
  const size_t lvl_stride = (size_t)n1l * n2 * sizeof(double), page = (size_t)2 << 20;
  const size_t remote_levels = (size_t)(ng3 - n3l);
  const size_t remote_pages = lvl_stride >= page ? remote_levels : (remote_levels * lvl_stride + page - 1) / page;
  const bool zcopy = sp->p2p && (zcopy_env >= 0 ? zcopy_env == 1 : remote_pages > 600);
  {
    ColGeom og;
    for (int q = 0; q < FB_MAX_RANKS; ++q) og.ptr[q] = nullptr;
    for (int q = 0; q < P; ++q) og.ptr[q] = sp->peer_recv[q];
    og.n3l = n3l; og.koff = (long)r * (long)chunk;
    StageTimer t(ST_Z);
    if (int rc = run_z(sp, (long)n1l * n2, ng3, sp->lam_win.as<double>(), W2, (sp->p2p && !zcopy) ? &og : nullptr, periodic, singular)) return rc;
    if (zcopy) {
      static cudaStream_t cs[FB_MAX_RANKS] = {};
      static cudaEvent_t ev_z = nullptr, ev_done[FB_MAX_RANKS];
      if (!ev_z) {
        CK(cudaEventCreateWithFlags(&ev_z, cudaEventDisableTiming));
        for (int q = 0; q < FB_MAX_RANKS; ++q) {
          CK(cudaStreamCreateWithFlags(&cs[q], cudaStreamNonBlocking));
          CK(cudaEventCreateWithFlags(&ev_done[q], cudaEventDisableTiming));
        }
      }
      CK(cudaEventRecord(ev_z, g_stream));
      for (int d = 0; d < P; ++d) {
        const int q = (r + d) % P;                       // start with the own chunk, then ring order: no two ranks hit one peer first
        CK(cudaStreamWaitEvent(cs[q], ev_z, 0));
        CK(cudaMemcpyAsync(sp->peer_recv[q] + (size_t)r * chunk, W2 + (size_t)q * chunk, chunk * sizeof(double),
                           cudaMemcpyDeviceToDevice, cs[q]));
        CK(cudaEventRecord(ev_done[q], cs[q]));
        CK(cudaStreamWaitEvent(g_stream, ev_done[q], 0));
      }
    }
  }
  {
    StageTimer t(ST_EXCH_B);
    if (sp->p2p) { if (int rc = p2p_barrier(sp)) return rc; }
    else if (g_a2a(g_a2a_ctx, W2, R, chunk * sizeof(double), (void*)g_stream)) return fail(FLUTAS_B200_ERR_CUDA, "all-to-all callback failed");
  }
  {
    SpecGeom sg;
    for (int q = 0; q < FB_MAX_RANKS; ++q) sg.ptr[q] = nullptr;
    for (int q = 0; q < P; ++q) sg.ptr[q] = R + (size_t)q * chunk;
    sg.n1l = n1l; sg.koff = 0;
    StageTimer t(ST_YB);
    if (int rc = run_y<false>(sp->py, W1, n1, n3l, sg)) return rc;
  }
  { StageTimer t(ST_XB); if (int rc = run_x<false>(sp->px, W1, gw, pd, gp, normfft)) return rc; }

  if (host_p) {
    CK(cudaMemcpyAsync(p, pd, pcount * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (sp->p2p && sp->h_err && *sp->h_err)
      return fail(FLUTAS_B200_ERR_CUDA, "solver_slab: cross-GPU barrier time-out (%d): a peer did not arrive; p is invalid", *sp->h_err);
  }
  return FLUTAS_B200_OK;
}

int flutas_b200_fillps(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                       const double* dzfi, double dti, double rho0, const double* u, const double* v,
                       const double* w, double* p) {
  (void)dzi;
  if (int rc = ensure_device()) return rc;
  if (!dzfi || !u || !v || !w || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "halo widths must be >= 1");
  StencilGeom g = stencil_geom(nx, ny, nz, nh_u);
  const size_t ucount = (size_t)g.su1 * g.su2 * (nz + 2 * nh_u), pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fu, fv, fw, fp;
  if (int rc = stage_in(fu, 0, u, ucount, true)) return rc;
  if (int rc = stage_in(fv, 1, v, ucount, true)) return rc;
  if (int rc = stage_in(fw, 2, w, ucount, true)) return rc;
  if (int rc = stage_in(fp, 3, p, pcount, true)) return rc;     // halos of p must survive
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzfi, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  static const bool vec_env = [] { const char* e = getenv("FLUTAS_B200_CORREC_VEC"); return !(e && e[0] == '0'); }();
  const bool vec2 = vec_env && (nx % 2 == 0) && (nh_u % 2 == 1) && !((uintptr_t)fu.dev % 16) && !((uintptr_t)fv.dev % 16) &&
                    !((uintptr_t)fw.dev % 16) && !((uintptr_t)fp.dev % 16);          // as in flutas_b200_correc
  {
    StageTimer t(ST_FILLPS);
    if (vec2) {
      dim3 g2((nx / 2 + 1 + 63) / 64, (ny + 3) / 4, nz);
      fillps_vec2_kernel<<<g2, blk, 0, g_stream>>>(g, dti * dxi, dti * dyi, dti, g_coef.as<double>() + (nh_d - 1), rho0,
                                                   fu.dev, fv.dev, fw.dev, fp.dev);
    } else {
      fillps_kernel<<<grd, blk, 0, g_stream>>>(g, dti * dxi, dti * dyi, dti, g_coef.as<double>() + (nh_d - 1), rho0,
                                               fu.dev, fv.dev, fw.dev, fp.dev);
    }
  }
  LAUNCHED();
  if (int rc = stage_out(fp)) return rc;
  if (fp.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

int flutas_b200_updt_rhs_b(int nx, int ny, int nz, const char cbc[6], const double* rhsbx, const double* rhsby,
                           const double* rhsbz, double* p) {
  if (int rc = ensure_device()) return rc;
  if (!cbc || !rhsbx || !rhsby || !rhsbz || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  StencilGeom g = stencil_geom(nx, ny, nz, 1);
  const size_t pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fp;
  if (int rc = stage_in(fp, 3, p, pcount, true)) return rc;
  const size_t cx = 2 * (size_t)ny * nz, cy = 2 * (size_t)nx * nz, cz = 2 * (size_t)nx * ny;
  if (int rc = g_coef.reserve((cx + cy + cz) * sizeof(double))) return rc;
  double* rx = g_coef.as<double>();
  double* ry = rx + cx;
  double* rz = ry + cy;
  // periodic directions contribute nothing (bc_rhs factor = 0, initsolver.f90:263-294); skipping the
  // launch there is a bit-exact no-op (+0.0).  Boundary arrays that already live on the device are used in place
  // (a device-resident time loop uploads them once); host arrays are staged, and only for the directions that need them.
  auto face_array = [&](const double* src, double* stage, size_t cnt, const double** out) -> int {
    if (on_device(src)) { *out = src; return 0; }
    CK(cudaMemcpyAsync(stage, src, cnt * sizeof(double), cudaMemcpyDefault, g_stream));
    *out = stage;
    return 0;
  };
  const double *dx_ = nullptr, *dy_ = nullptr, *dz_ = nullptr;
  if (cbc[0] != 'P' || cbc[1] != 'P') {
    if (int rc = face_array(rhsbx, rx, cx, &dx_)) return rc;
    updt_rhs_b_kernel<<<(unsigned)((cx / 2 + 255) / 256), 256, 0, g_stream>>>(g, dx_, fp.dev);
    LAUNCHED();
  }
  if (cbc[2] != 'P' || cbc[3] != 'P') {
    if (int rc = face_array(rhsby, ry, cy, &dy_)) return rc;
    updt_rhs_b_y_kernel<<<(unsigned)((cy / 2 + 255) / 256), 256, 0, g_stream>>>(g, dy_, fp.dev);
    LAUNCHED();
  }
  if (cbc[4] != 'P' || cbc[5] != 'P') {
    if (int rc = face_array(rhsbz, rz, cz, &dz_)) return rc;
    const double* rz = dz_;
    // z-slab decomposition (flutas_b200_init): the z faces belong to ranks 0 and nranks-1 only -- on the others the
    // neighbour is a rank, not MPI_PROC_NULL (bound.f90:915,929); x and y are never decomposed in this layout
    const int sides = (g_nranks == 1) ? 3 : ((g_rank == 0 ? 1 : 0) | (g_rank == g_nranks - 1 ? 2 : 0));
    if (sides) {
      updt_rhs_b_z_kernel<<<(unsigned)((cz / 2 + 255) / 256), 256, 0, g_stream>>>(g, rz, fp.dev, sides);
      LAUNCHED();
    }
  }
  if (int rc = stage_out(fp)) return rc;
  if (fp.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

int flutas_b200_correc(int nx, int ny, int nz, int nh_d, int nh_u, double dxi, double dyi, double dzi,
                       const double* dzci, double dt, double rho0, const double* p, double* u, double* v,
                       double* w, const double* rho) {
  (void)dzi; (void)rho;
  if (int rc = ensure_device()) return rc;
  if (!dzci || !u || !v || !w || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "halo widths must be >= 1");
  StencilGeom g = stencil_geom(nx, ny, nz, nh_u);
  const size_t ucount = (size_t)g.su1 * g.su2 * (nz + 2 * nh_u), pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fu, fv, fw, fp;
  if (int rc = stage_in(fu, 0, u, ucount, true)) return rc;
  if (int rc = stage_in(fv, 1, v, ucount, true)) return rc;
  if (int rc = stage_in(fw, 2, w, ucount, true)) return rc;
  if (int rc = stage_in(fp, 3, p, pcount, true)) return rc;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzci, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  const double rho0i = 1.0 / rho0;
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  // two points per thread with 16-byte accesses when every row start is 16-byte aligned (even nx, aligned bases; the pair
  // element index parity is that of nh_u - 1 for u,v,w and 0 for p: both even only for odd nh_u, i.e. nh_u = 1 or 3)
  static const bool vec_env = [] { const char* e = getenv("FLUTAS_B200_CORREC_VEC"); return !(e && e[0] == '0'); }();
  const bool vec2 = vec_env && (nx % 2 == 0) && (nh_u % 2 == 1) && !((uintptr_t)fu.dev % 16) && !((uintptr_t)fv.dev % 16) &&
                    !((uintptr_t)fw.dev % 16) && !((uintptr_t)fp.dev % 16);
  {
    StageTimer t(ST_CORREC);
    if (vec2) {
      dim3 g2((nx / 2 + 1 + 63) / 64, (ny + 3) / 4, nz);
      correc_vec2_kernel<<<g2, blk, 0, g_stream>>>(g, dt * dxi, dt * dyi, dt, g_coef.as<double>() + (nh_d - 1), rho0i,
                                                   fp.dev, fu.dev, fv.dev, fw.dev);
    } else {
      correc_kernel<<<grd, blk, 0, g_stream>>>(g, dt * dxi, dt * dyi, dt, g_coef.as<double>() + (nh_d - 1), rho0i,
                                               fp.dev, fu.dev, fv.dev, fw.dev);
    }
  }
  LAUNCHED();
  if (int rc = stage_out(fu)) return rc;
  if (int rc = stage_out(fv)) return rc;
  if (int rc = stage_out(fw)) return rc;
  if (fu.staged || fv.staged || fw.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

int flutas_b200_pres_sp_src(int nx, int ny, int nz, double f_t12, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                            const double* dzci, double rho0i, const double* pold, double* u, double* v, double* w) {
  (void)dzi;
  if (int rc = ensure_device()) return rc;
  if (!dzci || !u || !v || !w || !pold) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "halo widths must be >= 1");
  StencilGeom g = stencil_geom(nx, ny, nz, nh_u);
  const size_t ucount = (size_t)g.su1 * g.su2 * (nz + 2 * nh_u), pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fu, fv, fw, fp;
  if (int rc = stage_in(fu, 0, u, ucount, true)) return rc;
  if (int rc = stage_in(fv, 1, v, ucount, true)) return rc;
  if (int rc = stage_in(fw, 2, w, ucount, true)) return rc;
  if (int rc = stage_in(fp, 3, pold, pcount, true)) return rc;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzci, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  pres_sp_src_kernel<<<grd, blk, 0, g_stream>>>(g, f_t12, dxi, dyi, g_coef.as<double>() + (nh_d - 1), rho0i, fp.dev, fu.dev,
                                                fv.dev, fw.dev);
  LAUNCHED();
  if (int rc = stage_out(fu)) return rc;
  if (int rc = stage_out(fv)) return rc;
  if (int rc = stage_out(fw)) return rc;
  if (fu.staged || fv.staged || fw.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

int flutas_b200_pres_tw_src(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u, const double* dzci,
                            double rho0i, double f_t12, double f_t12_o, const double* p, const double* pold,
                            const double* rho, double* u, double* v, double* w) {
  (void)dzi;
  if (int rc = ensure_device()) return rc;
  if (!dzci || !u || !v || !w || !p || !pold || !rho) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "halo widths must be >= 1");
  if (f_t12_o == 0.0) return fail(FLUTAS_B200_ERR_ARG, "pres_tw_src: f_t12_o must not be zero");
  StencilGeom g = stencil_geom(nx, ny, nz, nh_u);
  const size_t ucount = (size_t)g.su1 * g.su2 * (nz + 2 * nh_u), pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fu, fv, fw, fp, fo, fr;
  if (int rc = stage_in(fu, 0, u, ucount, true)) return rc;
  if (int rc = stage_in(fv, 1, v, ucount, true)) return rc;
  if (int rc = stage_in(fw, 2, w, ucount, true)) return rc;
  if (int rc = stage_in(fp, 3, p, pcount, true)) return rc;
  if (int rc = stage_in(fo, 4, pold, pcount, true)) return rc;
  if (int rc = stage_in(fr, 5, rho, pcount, true)) return rc;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzci, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  const double f1 = 1.0 + (f_t12 / f_t12_o), f2 = (f_t12 / f_t12_o);      // source.f90:266-269
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  pres_tw_src_kernel<<<grd, blk, 0, g_stream>>>(g, dxi, dyi, g_coef.as<double>() + (nh_d - 1), rho0i, f_t12, f1, f2, fp.dev,
                                                fo.dev, fr.dev, fu.dev, fv.dev, fw.dev);
  LAUNCHED();
  if (int rc = stage_out(fu)) return rc;
  if (int rc = stage_out(fv)) return rc;
  if (int rc = stage_out(fw)) return rc;
  if (fu.staged || fv.staged || fw.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

int flutas_b200_pold_update(int nx, int ny, int nz, int mode, double* p, double* pold) {
  if (int rc = ensure_device()) return rc;
  if (!p || !pold) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (mode != 0 && mode != 1) return fail(FLUTAS_B200_ERR_ARG, "pold_update: mode must be 0 (pold = p) or 1 (p = pold + p)");
  StencilGeom g = stencil_geom(nx, ny, nz, 1);
  const size_t pcount = (size_t)g.sp1 * g.sp2 * (nz + 2);
  FieldRef fp, fo;
  if (int rc = stage_in(fp, 0, p, pcount, true)) return rc;
  if (int rc = stage_in(fo, 1, pold, pcount, true)) return rc;
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  pold_update_kernel<<<grd, blk, 0, g_stream>>>(g, mode, fp.dev, fo.dev);
  LAUNCHED();
  if (mode == 0) { if (int rc = stage_out(fo)) return rc; }
  else { if (int rc = stage_out(fp)) return rc; }
  if (fp.staged || fo.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

// load(io,filename,n,fld), src/load.f90:21-89: restart fields are headerless raw FP64 in GLOBAL column-major (ng1,ng2,ng3)
// order, independent of the decomposition (each rank reads/writes its block through a subarray file view,
// src/2decomp/io_write_var.f90:28-60).  Here: pread/pwrite of the block's x-rows (one contiguous run for a z-slab),
// staged through host memory; `fld` may be a host or a device pointer with `nh` halo cells on every side.
int flutas_b200_load(char io, const char* filename, const int ng[3], const int n[3], const int start[3], int nh, double* fld) {
  if (!filename || !ng || !n || !start || !fld) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (io != 'r' && io != 'w') return fail(FLUTAS_B200_ERR_ARG, "load: io must be 'r' or 'w'");
  if (nh < 0) return fail(FLUTAS_B200_ERR_ARG, "load: nh must be >= 0");
  for (int d = 0; d < 3; ++d)
    if (n[d] < 1 || start[d] < 0 || start[d] + n[d] > ng[d]) return fail(FLUTAS_B200_ERR_ARG, "load: block outside the global grid");
  const long long good = 8LL * ng[0] * ng[1] * ng[2];
  const size_t count = (size_t)n[0] * n[1] * n[2];
  const bool dev = on_device(fld);
  if (dev) { if (int rc = ensure_device()) return rc; }
  std::vector<double> dense;
  double* blk = fld;                                        // dense (n1,n2,n3) image of the block in host memory
  if (dev || nh > 0) { dense.resize(count); blk = dense.data(); }
  const size_t s1 = (size_t)n[0] + 2 * nh, s2 = (size_t)n[1] + 2 * nh;
  auto halo_copy = [&](bool to_fld) -> int {                // dense block <-> interior of the (possibly halo'd, possibly device) fld
    if (!dev && nh == 0) return 0;
    cudaMemcpy3DParms pm = {};
    cudaPitchedPtr pd = make_cudaPitchedPtr(blk, (size_t)n[0] * 8, (size_t)n[0], (size_t)n[1]);
    cudaPitchedPtr pf = make_cudaPitchedPtr(fld, s1 * 8, s1, s2);
    pm.extent = make_cudaExtent((size_t)n[0] * 8, (size_t)n[1], (size_t)n[2]);
    pm.kind = cudaMemcpyDefault;
    if (to_fld) { pm.srcPtr = pd; pm.dstPtr = pf; pm.dstPos = make_cudaPos((size_t)nh * 8, (size_t)nh, (size_t)nh); }
    else { pm.srcPtr = pf; pm.srcPos = make_cudaPos((size_t)nh * 8, (size_t)nh, (size_t)nh); pm.dstPtr = pd; }
    if (dev) { CK(cudaMemcpy3D(&pm)); return 0; }
    for (int k = 0; k < n[2]; ++k)                          // host array with halos: no CUDA call needed
      for (int j = 0; j < n[1]; ++j) {
        double* f = fld + (size_t)nh + s1 * ((size_t)(j + nh) + s2 * (size_t)(k + nh));
        double* b = blk + (size_t)n[0] * ((size_t)j + (size_t)n[1] * k);
        if (to_fld) memcpy(f, b, (size_t)n[0] * 8); else memcpy(b, f, (size_t)n[0] * 8);
      }
    return 0;
  };
  int fd = -1;
  if (io == 'r') {
    fd = open(filename, O_RDONLY);
    if (fd < 0) return fail(FLUTAS_B200_ERR_ARG, "load: the restarting field %s does not exist", filename);   // load.f90:40-46
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (long long)sb.st_size != good) {                                                 // load.f90:52-66
      const long long have = (long long)sb.st_size;
      close(fd);
      return fail(FLUTAS_B200_ERR_ARG, "load: checkpoint file %s has incorrect size (expected %lld, actual %lld)", filename, good, have);
    }
  } else {
    if (dev) CK(cudaStreamSynchronize(g_stream));
    if (int rc = halo_copy(false)) return rc;
    fd = open(filename, O_CREAT | O_WRONLY, 0644);
    if (fd < 0) return fail(FLUTAS_B200_ERR_ARG, "load: cannot create %s", filename);
    if (ftruncate(fd, (off_t)good) != 0) { close(fd); return fail(FLUTAS_B200_ERR_ARG, "load: cannot size %s", filename); }
  }
  const bool rows_contiguous = (n[0] == ng[0]) && (n[1] == ng[1]);
  const size_t run = rows_contiguous ? count : (size_t)n[0];           // doubles per contiguous file run
  const size_t nruns = count / run;
  for (size_t r = 0; r < nruns; ++r) {
    const size_t j = rows_contiguous ? 0 : r % n[1], k = rows_contiguous ? 0 : r / n[1];
    const off_t off = 8 * ((off_t)start[0] + (off_t)ng[0] * ((off_t)(start[1] + j) + (off_t)ng[1] * (off_t)(start[2] + k)));
    char* q = reinterpret_cast<char*>(blk + r * run);
    size_t left = run * 8, done_b = 0;
    while (left) {
      const ssize_t got = (io == 'r') ? pread(fd, q + done_b, left, off + (off_t)done_b) : pwrite(fd, q + done_b, left, off + (off_t)done_b);
      if (got <= 0) { close(fd); return fail(FLUTAS_B200_ERR_ARG, "load: short %s on %s", io == 'r' ? "read" : "write", filename); }
      left -= (size_t)got; done_b += (size_t)got;
    }
  }
  close(fd);
  if (io == 'r') {
    if (dev) CK(cudaStreamSynchronize(g_stream));          // kernels still reading the old contents of fld on the library stream
    if (int rc = halo_copy(true)) return rc;
  }
  return FLUTAS_B200_OK;
}

int flutas_b200_set_halo_exchange(flutas_b200_halo_fn fn, void* ctx) {
  g_halo = fn;
  g_halo_ctx = ctx;
  return FLUTAS_B200_OK;
}

int flutas_b200_boundp(const char cbc[6], const int n[3], const double bc[6], int nh_d, int nh_p, const double dl[3],
                       const double* dzc, const double* dzf, double* p) {
  (void)dzf;
  if (int rc = ensure_device()) return rc;
  if (!cbc || !n || !bc || !dl || !dzc || !p) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_p != 1) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "boundp: only nh_p = 1 (pressure halo) is on this path");
  if (nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "nh_d must be >= 1");
  for (int q = 0; q < 6; ++q)
    if (cbc[q] != 'P' && cbc[q] != 'D' && cbc[q] != 'N') return fail(FLUTAS_B200_ERR_ARG, "bad boundary type '%c'", cbc[q]);
  const int nx = n[0], ny = n[1], nz = n[2];
  const long s1 = nx + 2, s2 = ny + 2, s3 = nz + 2;
  const size_t pcount = (size_t)s1 * s2 * s3;
  FieldRef fp;
  if (int rc = stage_in(fp, 3, p, pcount, true)) return rc;
  double* pd = fp.dev;
  // z metric at the two walls: dzc(0), dzc(nz) (bound.f90:209-218); only needed for a non-zero Neumann value
  double dzc_lo = 0.0, dzc_hi = 0.0;
  if ((cbc[4] == 'N' && bc[4] != 0.0) || (cbc[5] == 'N' && bc[5] != 0.0)) {
    CK(cudaMemcpyAsync(&dzc_lo, dzc + (nh_d - 1), sizeof(double), cudaMemcpyDefault, g_stream));
    CK(cudaMemcpyAsync(&dzc_hi, dzc + (nh_d - 1) + nz, sizeof(double), cudaMemcpyDefault, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  const FaceGeom gx{1, s1, s1 * s2, nx, (int)s2, (int)s3};
  const FaceGeom gy{s1, 1, s1 * s2, ny, (int)s1, (int)s3};
  const FaceGeom gz{s1 * s2, 1, s1, nz, (int)s1, (int)s2};
  auto nblk = [](const FaceGeom& g) { return (unsigned)(((long)g.na * g.nb + 255) / 256); };
  auto wrap = [&](const FaceGeom& g) -> int {
    boundp_wrap_kernel<<<nblk(g), 256, 0, g_stream>>>(g, pd);
    LAUNCHED();
    return 0;
  };
  // set_bc for one side (bound.f90:247-268): D (centred): 2 v - p_in ; N: p_in -/+ dr v
  auto face = [&](const FaceGeom& g, int side, char type, double value, double dr) -> int {
    double factor = value, sgn = 0.0;
    if (type == 'D') { factor = 2.0 * factor; sgn = -1.0; }
    if (type == 'N') { factor = (side == 0) ? -dr * factor : dr * factor; sgn = 1.0; }
    boundp_face_kernel<<<nblk(g), 256, 0, g_stream>>>(g, side, factor, sgn, pd);
    LAUNCHED();
    return 0;
  };
  const bool py = (cbc[2] == 'P' && cbc[3] == 'P'), pz = (cbc[4] == 'P' && cbc[5] == 'P');
  const int P = g_nranks, r = g_rank;
  // 1. updthalo along y (the rank is its own y neighbour in the z-slab layout), bound.f90:182
  if (py) { if (int rc = wrap(gy)) return rc; }
  // 2. updthalo along z, bound.f90:183
  int lo = -1, hi = -1;                                   // bottom / top neighbours (MPI_CART_SHIFT, initmpi.f90:127)
  if (P > 1) {
    lo = (r > 0) ? r - 1 : (pz ? P - 1 : -1);
    hi = (r < P - 1) ? r + 1 : (pz ? 0 : -1);
  }
  if (P == 1) {
    if (pz) { if (int rc = wrap(gz)) return rc; }
  } else {
    if (!g_halo) return fail(FLUTAS_B200_ERR_ARG, "boundp on %d ranks needs flutas_b200_set_halo_exchange", P);
    const size_t plane = (size_t)s1 * s2;
    if (g_halo(g_halo_ctx, pd + plane, pd + plane * nz, pd, pd + plane * (nz + 1), plane, lo, hi, (void*)g_stream))
      return fail(FLUTAS_B200_ERR_CUDA, "halo exchange callback failed");
  }
  // 3.-5. set_bc on the faces this rank owns (x is never decomposed: left = right = MPI_PROC_NULL, initmpi.f90:124)
  if (cbc[0] == 'P') { if (int rc = wrap(gx)) return rc; }
  else {
    if (int rc = face(gx, 0, cbc[0], bc[0], dl[0])) return rc;
    if (int rc = face(gx, 1, cbc[1], bc[1], dl[0])) return rc;
  }
  if (!py) {
    if (int rc = face(gy, 0, cbc[2], bc[2], dl[1])) return rc;
    if (int rc = face(gy, 1, cbc[3], bc[3], dl[1])) return rc;
  }
  if (!pz) {
    if (lo < 0 && (P == 1 || r == 0)) { if (int rc = face(gz, 0, cbc[4], bc[4], dzc_lo)) return rc; }
    if (hi < 0 && (P == 1 || r == P - 1)) { if (int rc = face(gz, 1, cbc[5], bc[5], dzc_hi)) return rc; }
  }
  if (int rc = stage_out(fp)) return rc;
  if (fp.staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

// bounduvw(cbc,n,bc,nh_d,nh_u,halo,isoutflow,dl,dzc,dzf,u,v,w), src/bound.f90:17-144, in the reference's step order:
// updthalo along y and z for u, v, w (:55-62), set_bc on the x faces, the y faces, the z faces this rank owns (:68-128),
// then outflow (:131-141).  z-slab decomposition: the z halo layers (nh_u planes per side) go through the callback of
// flutas_b200_set_halo_exchange; x and y are never decomposed in this layout.
int flutas_b200_bounduvw(const char cbc[18], const int n[3], const double bc[18], int nh_d, int nh_u, const int isoutflow[6],
                         const double dl[3], const double* dzc, const double* dzf, double* u, double* v, double* w) {
  if (int rc = ensure_device()) return rc;
  if (!cbc || !n || !bc || !isoutflow || !dl || !dzc || !dzf || !u || !v || !w) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_u > FB_MAX_HALO) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "bounduvw: halo width %d (1..%d)", nh_u, (int)FB_MAX_HALO);
  if (nh_d < nh_u) return fail(FLUTAS_B200_ERR_ARG, "bounduvw: nh_d = %d < nh_u = %d", nh_d, nh_u);
  for (int q = 0; q < 18; ++q)
    if (cbc[q] != 'P' && cbc[q] != 'D' && cbc[q] != 'N') return fail(FLUTAS_B200_ERR_ARG, "bad boundary type '%c'", cbc[q]);
  const int nx = n[0], ny = n[1], nz = n[2], nh = nh_u;
  if (nx < nh + 1 || ny < nh + 1 || nz < nh + 1) return fail(FLUTAS_B200_ERR_UNSUPPORTED, "bounduvw: grid smaller than the halo");
  const HaloField g = halo_field(nx, ny, nz, nh);
  const size_t count = (size_t)g.s[2] * (nz + 2 * nh);
  FieldRef f[3];
  double* in[3] = {u, v, w};
  for (int q = 0; q < 3; ++q)
    if (int rc = stage_in(f[q], q, in[q], count, true)) return rc;
  auto C = [&](int ib, int idir, int fld) { return cbc[ib + 2 * (idir + 3 * fld)]; };          // cbc(0:1,3,3), Fortran order
  auto B = [&](int ib, int idir, int fld) { return bc[ib + 2 * (idir + 3 * fld)]; };
  auto periodic = [&](int idir) {
    for (int ib = 0; ib < 2; ++ib) for (int fld = 0; fld < 3; ++fld) if (C(ib, idir, fld) != 'P') return false;
    return true;
  };
  const bool py = periodic(1), pz = periodic(2);
  const int P = g_nranks, r = g_rank;
  auto launch_bc = [&](double* p, int idir, int ib, int mode, double sgn, const BcFactor& fac) -> int {
    const int da = (idir == 0) ? 1 : 0, db = (idir == 2) ? 1 : 2;
    const long cnt = (long)(g.n[da] + 2 * nh) * (g.n[db] + 2 * nh);
    set_bc_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(p, g, idir, ib, mode, sgn, fac);
    LAUNCHED();
    return 0;
  };
  // set_bc (:238-266): factor and sign from the boundary type; dr[q] only matters for Neumann
  auto set_bc = [&](double* p, char type, int ib, int idir, bool centered, double value, const double* dr) -> int {
    BcFactor fac;
    double sgn = 0.0;
    for (int q = 0; q < nh; ++q) fac.f[q] = value;
    if (type == 'P') return launch_bc(p, idir, ib, BC_WRAP, 0.0, fac);
    if (type == 'D' && centered) { for (int q = 0; q < nh; ++q) fac.f[q] = 2.0 * fac.f[q]; sgn = -1.0; }
    if (type == 'N') { for (int q = 0; q < nh; ++q) fac.f[q] = (ib == 0) ? -dr[q] * fac.f[q] : dr[q] * fac.f[q]; sgn = 1.0; }
    const int mode = centered ? BC_CENTERED : (type == 'D') ? BC_FACE_D : BC_FACE_N;
    return launch_bc(p, idir, ib, mode, sgn, fac);
  };
  // 1. updthalo along y, then z, field by field (:55-62)
  const BcFactor none{};
  const size_t plane = (size_t)g.s[2];
  int lo = -1, hi = -1;
  if (P > 1) { lo = (r > 0) ? r - 1 : (pz ? P - 1 : -1); hi = (r < P - 1) ? r + 1 : (pz ? 0 : -1); }
  if (P > 1 && !g_halo) return fail(FLUTAS_B200_ERR_ARG, "bounduvw on %d ranks needs flutas_b200_set_halo_exchange", P);
  for (int q = 0; q < 3; ++q) {
    double* p = f[q].dev;
    if (py) { if (int rc = launch_bc(p, 1, 0, BC_WRAP, 0.0, none)) return rc; }
    if (P == 1) { if (pz) { if (int rc = launch_bc(p, 2, 0, BC_WRAP, 0.0, none)) return rc; } }
    else if (g_halo(g_halo_ctx, p + plane * nh, p + plane * nz, p, p + plane * (nz + nh), plane * nh, lo, hi, (void*)g_stream))
      return fail(FLUTAS_B200_ERR_CUDA, "halo exchange callback failed");
  }
  // z metrics at the walls for Neumann values: dr(q) = dzc(-q) / dzf(-q) / dzc(n3+q) / dzf(n3+q) (:95-119)
  std::vector<double> hzc, hzf;
  bool need_dr = false;
  if (!pz) for (int ib = 0; ib < 2; ++ib) for (int fld = 0; fld < 3; ++fld) if (C(ib, 2, fld) == 'N' && B(ib, 2, fld) != 0.0) need_dr = true;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (need_dr) {
    hzc.resize(nd); hzf.resize(nd);
    CK(cudaMemcpyAsync(hzc.data(), dzc, nd * sizeof(double), cudaMemcpyDefault, g_stream));
    CK(cudaMemcpyAsync(hzf.data(), dzf, nd * sizeof(double), cudaMemcpyDefault, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  double dr[FB_MAX_HALO];
  // 2. x faces (left = right = MPI_PROC_NULL in this layout), 3. y faces unless periodic, 4. z faces this rank owns
  for (int idir = 0; idir < 3; ++idir) {
    if ((idir == 1 && py) || (idir == 2 && pz)) continue;
    for (int ib = 0; ib < 2; ++ib) {
      if (idir == 2 && P > 1 && ((ib == 0 && r != 0) || (ib == 1 && r != P - 1))) continue;
      for (int fld = 0; fld < 3; ++fld) {
        const bool centered = (fld != idir);
        for (int q = 0; q < nh; ++q) {
          if (idir < 2) dr[q] = dl[idir];
          else if (!need_dr) dr[q] = 1.0;                   // multiplies a zero value (or is unused)
          else dr[q] = (centered ? hzc : hzf)[(size_t)((ib == 0 ? -q : nz + q) + nh_d - 1)];
        }
        if (int rc = set_bc(f[fld].dev, C(ib, idir, fld), ib, idir, centered, B(ib, idir, fld), dr)) return rc;
      }
    }
  }
  // 5. outflow (:131-141)
  bool any_out = false;
  for (int q = 0; q < 6; ++q) any_out = any_out || (isoutflow[q] != 0);
  if (any_out) {
    const double* dzf_dev = dzf;
    if (!on_device(dzf)) {
      if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
      CK(cudaMemcpyAsync(g_coef.p, dzf, nd * sizeof(double), cudaMemcpyHostToDevice, g_stream));
      dzf_dev = g_coef.as<double>();
    }
    for (int q = 0; q < 3; ++q)
      for (int ib = 0; ib < 2; ++ib) {
        if (!isoutflow[ib + 2 * q]) continue;
        if (q == 2 && P > 1 && ((ib == 0 && r != 0) || (ib == 1 && r != P - 1))) continue;   // top / bottom == MPI_PROC_NULL
        if ((q == 1 && py) || (q == 2 && pz)) continue;                                          // neighbour is a rank, not a wall
        const int da = (q == 0) ? 1 : 0, db = (q == 2) ? 1 : 2;
        const long cnt = (long)g.n[da] * g.n[db];
        outflow_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, g_stream>>>(g, (q + 1) * (ib == 0 ? -1 : 1), nh_d, dl[0], dl[1], dzf_dev,
                                                                           f[0].dev, f[1].dev, f[2].dev);
        LAUNCHED();
      }
  }
  for (int q = 0; q < 3; ++q) if (int rc = stage_out(f[q])) return rc;
  if (f[0].staged || f[1].staged || f[2].staged) CK(cudaStreamSynchronize(g_stream));
  return FLUTAS_B200_OK;
}

// The field reduction of chkdt_sp / chkdt_tw (src/chkdt.f90:62-85 = :150-173): this rank's max over cells of the three
// convective inverse time scales.  The caller all-reduces (MPI_MAX, :92,183) and applies the scalar formulas (:93-110,
// :184-196: `if(dti.eq.0) dti = 1`, viscous / gravity limits, dtmax) exactly as the reference does.  Synchronous.
int flutas_b200_chkdt(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u, const double* dzci,
                      const double* dzfi, const double* u, const double* v, const double* w, double* dti) {
  (void)dzi;
  if (int rc = ensure_device()) return rc;
  if (!dzci || !dzfi || !u || !v || !w || !dti) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  if (nh_u < 1 || nh_d < 1) return fail(FLUTAS_B200_ERR_ARG, "halo widths must be >= 1");
  const HaloField g = halo_field(nx, ny, nz, nh_u);
  const size_t count = (size_t)g.s[2] * (nz + 2 * nh_u);
  FieldRef fu, fv, fw;
  if (int rc = stage_in(fu, 0, u, count, true)) return rc;
  if (int rc = stage_in(fv, 1, v, count, true)) return rc;
  if (int rc = stage_in(fw, 2, w, count, true)) return rc;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(2 * nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzci, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  CK(cudaMemcpyAsync(g_coef.as<double>() + nd, dzfi, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  const long nparts = (long)grd.x * grd.y * grd.z;
  if (int rc = g_red.reserve(((size_t)nparts + 2) * sizeof(double))) return rc;
  double* part = g_red.as<double>();
  chkdt_kernel<<<grd, blk, 0, g_stream>>>(g, nh_d, dxi, dyi, g_coef.as<double>(), g_coef.as<double>() + nd, fu.dev, fv.dev, fw.dev, part);
  LAUNCHED();
  max_final_kernel<<<1, 256, 0, g_stream>>>(nparts, part, part + nparts);
  LAUNCHED();
  double res = 0.0;
  CK(cudaMemcpyAsync(&res, part + nparts, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  *dti = res;
  return FLUTAS_B200_OK;
}

int flutas_b200_chkdiv(int nx, int ny, int nz, double dxi, double dyi, double dzi, int nh_d, int nh_u,
                       const double* dzfi, const double* u, const double* v, const double* w, double* divtot,
                       double* divmax) {
  (void)dzi;
  if (int rc = ensure_device()) return rc;
  if (!dzfi || !u || !v || !w || !divtot || !divmax) return fail(FLUTAS_B200_ERR_ARG, "null argument");
  StencilGeom g = stencil_geom(nx, ny, nz, nh_u);
  const size_t ucount = (size_t)g.su1 * g.su2 * (nz + 2 * nh_u);
  FieldRef fu, fv, fw;
  if (int rc = stage_in(fu, 0, u, ucount, true)) return rc;
  if (int rc = stage_in(fv, 1, v, ucount, true)) return rc;
  if (int rc = stage_in(fw, 2, w, ucount, true)) return rc;
  const size_t nd = (size_t)nz + 2 * nh_d;
  if (int rc = g_coef.reserve(nd * sizeof(double))) return rc;
  CK(cudaMemcpyAsync(g_coef.p, dzfi, nd * sizeof(double), cudaMemcpyDefault, g_stream));
  dim3 blk(64, 4, 1), grd((nx + 63) / 64, (ny + 3) / 4, nz);
  const long nparts = (long)grd.x * grd.y * grd.z;
  if (int rc = g_red.reserve((2 * (size_t)nparts + 2) * sizeof(double))) return rc;
  double* ps = g_red.as<double>();
  double* pm = ps + nparts;
  double* out = pm + nparts;
  chkdiv_kernel<<<grd, blk, 0, g_stream>>>(g, dxi, dyi, g_coef.as<double>() + (nh_d - 1), fu.dev, fv.dev, fw.dev, ps, pm);
  LAUNCHED();
  chkdiv_final_kernel<<<1, 256, 0, g_stream>>>(nparts, ps, pm, out);
  LAUNCHED();
  double res[2];
  CK(cudaMemcpyAsync(res, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  *divtot = res[0];
  *divmax = res[1];
  return FLUTAS_B200_OK;
}

}  // extern "C"
