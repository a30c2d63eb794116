// libflutas_b200_fftw.so -- the one C-ABI seam the reference already has on this path: FFTW's.
//
// src/fft.f90 plans through `fftw_plan_guru_r2r` (bind(C), interface src/fftw.f90:15-36, called at src/fft.f90:85-86,123-124)
// and executes / destroys through FFTW's legacy Fortran entry points, called without an interface and therefore
// compiler-mangled with every argument by reference: `dfftw_execute_r2r(plan,arr,arr)` (:188-190), `dfftw_destroy_plan`
// (:165-175), `dfftw_init_threads`, `dfftw_plan_with_nthreads`, `dfftw_cleanup_threads` (:53-57,171-175, OpenMP builds).
// Linking this library in place of -lfftw3 lets fft.f90 / fftw.f90 / solver_cpu.f90 compile and run literally unchanged,
// with every transform on the GPU (host arrays are staged per call: this is the compatibility route, the fused route is
// the shim's solver_cpu -> flutas_b200_solver).  It is a separate library so that libflutas_b200.so never shadows a real
// FFTW that other parts of a host program may link.
//
// Errors: FFTW returns a null plan and the reference ignores it (istat unused, src/fft.f90:92-146); here a failure
// prints flutas_b200_last_error() and exits -- never silently wrong data.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/flutas_b200.h"

namespace {
[[noreturn]] void die(const char* where) {
  std::fprintf(stderr, "flutas_b200 (%s): %s\n", where, flutas_b200_last_error());
  std::exit(1);
}
}  // namespace

extern "C" {

struct fftw_iodim { int n, is, os; };          // type, bind(C) :: fftw_iodim, src/fftw.f90:11-13

void* fftw_plan_guru_r2r(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims,
                         double* in, double* out, const int* kind, unsigned flags) {
  (void)flags;
  if (rank != 1 || howmany_rank != 2 || !dims || !howmany_dims || !kind || in != out || dims[0].is != dims[0].os ||
      howmany_dims[0].is != howmany_dims[0].os || howmany_dims[1].is != howmany_dims[1].os) {
    std::fprintf(stderr, "flutas_b200 (fftw_plan_guru_r2r): only the in-place rank-1 x 2-howmany plans of src/fft.f90:75-86,113-124\n");
    std::exit(1);
  }
  const int hn[2] = {howmany_dims[0].n, howmany_dims[1].n}, hs[2] = {howmany_dims[0].is, howmany_dims[1].is};
  void* plan = nullptr;
  if (flutas_b200_plan_r2r(dims[0].n, dims[0].is, hn, hs, kind[0], &plan) != FLUTAS_B200_OK) die("fftw_plan_guru_r2r");
  return plan;
}

// call dfftw_execute_r2r(plan, in, out): the plan variable itself is passed by reference
static void execute(void** plan, double* in, double* out) {
  int n[3];
  if (!plan || flutas_b200_plan_dims(*plan, n) != FLUTAS_B200_OK) die("dfftw_execute_r2r");
  if (in != out && flutas_b200_memcpy(out, in, sizeof(double) * (size_t)n[0] * n[1] * n[2]) != FLUTAS_B200_OK) die("dfftw_execute_r2r");
  if (flutas_b200_fft(*plan, n, out) != FLUTAS_B200_OK) die("dfftw_execute_r2r");
  if (flutas_b200_synchronize() != FLUTAS_B200_OK) die("dfftw_execute_r2r");      // FFTW's execute is synchronous
}
static void destroy(void** plan) {
  if (plan && *plan && flutas_b200_destroy_plan(*plan) != FLUTAS_B200_OK) die("dfftw_destroy_plan");
  if (plan) *plan = nullptr;
}

// gfortran / nvfortran / ifx mangling (trailing underscore) and the bare names
void dfftw_execute_r2r_(void** plan, double* in, double* out) { execute(plan, in, out); }
void dfftw_execute_r2r(void** plan, double* in, double* out) { execute(plan, in, out); }
void dfftw_destroy_plan_(void** plan) { destroy(plan); }
void dfftw_destroy_plan(void** plan) { destroy(plan); }
void dfftw_init_threads_(int* ierr) { if (ierr) *ierr = 1; }        // FFTW: non-zero = success
void dfftw_init_threads(int* ierr) { if (ierr) *ierr = 1; }
void dfftw_plan_with_nthreads_(const int* nthreads) { (void)nthreads; }
void dfftw_plan_with_nthreads(const int* nthreads) { (void)nthreads; }
void dfftw_cleanup_threads_(int* ierr) { (void)ierr; }
void dfftw_cleanup_threads(int* ierr) { (void)ierr; }

}  // extern "C"
