// TEST INFRASTRUCTURE ONLY. Single-threaded CPU emulation of the svd / eig kernel bodies in
// raven_b200/csrc/nxc_linalg3.cuh (the same source the GPU kernels instantiate), so that their
// control flow and numerics are exercised in the CPU test suite. Built by tests/test_la3_emu.py
// with g++; never linked into libnxcuda.so, never reachable from the product path.
#define __host__
#define __device__
#define __forceinline__ inline
#include <stdlib.h>
#include <string.h>
#include "../../raven_b200/csrc/nxc_linalg3.cuh"

template <class T>
static int svd_run(const T *a, int64_t m, int64_t n, int64_t ucols, int64_t vrows, T *u, double *s, T *vh) {
  const bool tall = m >= n;
  const int64_t pr = tall ? m : n, pc = tall ? n : m, ncu = tall ? ucols : vrows;
  T *gt = (T *)calloc((size_t)(pc * pr + 1), sizeof(T));
  T *wt = (T *)calloc((size_t)(pc * pc + 1), sizeof(T));
  T *ut = (T *)calloc((size_t)(ncu * pr + 1), sizeof(T));
  double *sg = (double *)calloc((size_t)pc + 1, sizeof(double));
  int *rk = (int *)calloc((size_t)pc + 1, sizeof(int));
  double *rown = (double *)calloc((size_t)pr + 1, sizeof(double));
  Cd *coef = (Cd *)calloc((size_t)pr + 1, sizeof(Cd));
  double red[1];
  int flags[2] = {0, 0}, status = 0;
  if (tall) {
    for (int64_t j = 0; j < n; j++)
      for (int64_t r = 0; r < m; r++) gt[j * m + r] = a[r * n + j];
  } else {
    memcpy(gt, a, (size_t)(m * n) * sizeof(T));
  }
  La3Thr t = {0, 1, 0, 1, 0, 1};
  la3_svd_body<T>(t, gt, wt, ut, u, vh, s, sg, rk, rown, coef, red, flags, m, n, ucols, vrows, 60, &status);
  free(gt); free(wt); free(ut); free(sg); free(rk); free(rown); free(coef);
  return status;
}

extern "C" int la3_emu_svd(int cdt, const void *a, int64_t m, int64_t n, int64_t ucols, int64_t vrows, void *u, double *s,
                           void *vh) {
  switch (cdt) {
    case 0: return svd_run<float>((const float *)a, m, n, ucols, vrows, (float *)u, s, (float *)vh);
    case 1: return svd_run<double>((const double *)a, m, n, ucols, vrows, (double *)u, s, (double *)vh);
    case 2: return svd_run<La3C32>((const La3C32 *)a, m, n, ucols, vrows, (La3C32 *)u, s, (La3C32 *)vh);
    default: return svd_run<La3C64>((const La3C64 *)a, m, n, ucols, vrows, (La3C64 *)u, s, (La3C64 *)vh);
  }
}

extern "C" int la3_emu_eig(const double *a, int64_t n, int vectors, double *w, double *v) {
  Cd *h = (Cd *)calloc((size_t)(n * n + 1), sizeof(Cd));
  Cd *z = (Cd *)calloc((size_t)(n * n + 1), sizeof(Cd));
  Cd *x = (Cd *)calloc((size_t)(n * n + 1), sizeof(Cd));
  Cd *vs = (Cd *)calloc((size_t)n + 1, sizeof(Cd));
  Cd *rs = (Cd *)calloc((size_t)n + 1, sizeof(Cd));
  double *rc = (double *)calloc((size_t)n + 1, sizeof(double));
  double *bal = (double *)calloc((size_t)n + 1, sizeof(double));
  int *flag = (int *)calloc((size_t)n + 1, sizeof(int));
  double red[1];
  int status = 0;
  memcpy(h, a, (size_t)(n * n) * sizeof(Cd));
  La3Thr t = {0, 1, 0, 1, 0, 1};
  la3_eig_body(t, h, z, x, (Cd *)v, (Cd *)w, vs, rc, rs, flag, bal, red, n, vectors, &status);
  free(h); free(z); free(x); free(vs); free(rs); free(rc); free(bal); free(flag);
  return status;
}

template <class T>
static int eigh_run(const T *a, int64_t n, int vectors, double *w, T *v) {
  T *gt = (T *)calloc((size_t)(n * n + 1), sizeof(T));
  T *vt = (T *)calloc((size_t)(n * n + 1), sizeof(T));
  double *sg = (double *)calloc((size_t)n + 1, sizeof(double));
  int *rk = (int *)calloc((size_t)n + 1, sizeof(int));
  double red[1];
  int flags[2] = {0, 0}, status = 0;
  La3Thr t = {0, 1, 0, 1, 0, 1};
  la3_eigh_body<T>(t, a, gt, vt, v, w, sg, rk, red, flags, n, vectors, 60, &status);
  free(gt); free(vt); free(sg); free(rk);
  return status;
}

extern "C" int la3_emu_eigh(int cdt, const void *a, int64_t n, int vectors, double *w, void *v) {
  switch (cdt) {
    case 0: return eigh_run<float>((const float *)a, n, vectors, w, (float *)v);
    case 1: return eigh_run<double>((const double *)a, n, vectors, w, (double *)v);
    case 2: return eigh_run<La3C32>((const La3C32 *)a, n, vectors, w, (La3C32 *)v);
    default: return eigh_run<La3C64>((const La3C64 *)a, n, vectors, w, (La3C64 *)v);
  }
}
