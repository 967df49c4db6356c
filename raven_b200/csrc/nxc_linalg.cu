// nxc_linalg.cu -- linalg tier 1: cholesky, triangular_solve, qr.
// Replaces caml_nx_c_cholesky / caml_nx_c_triangular_solve (reference: nx_c_tri.c:570-620, drivers
// :296-566) and caml_nx_c_qr (reference: nx_c_qr.c:421-433, driver :321-418).
//
// Contract kept from the reference: compute type per storage dtype (f16/bf16/fp8/f32 -> f32, f64,
// c32, c64; anything else is Invalid_argument "linalg requires a float or complex dtype",
// nx_c_linalg.h:183-200); batched over the leading dims, operands at arbitrary strides; cholesky
// reads the LOWER triangle only, fails with "matrix is not positive definite" on a pivot that is
// not > 0 (NaN included) and writes zeros in the other triangle, ~upper returns L^H;
// triangular_solve's `transpose` is the CONJUGATE transpose, an exactly-zero pivot (non-unit
// diagonal) is "triangular matrix is singular"; qr is Householder with LAPACK's pivot sign
// (beta = -sign(re alpha) * norm), tau = 0 for a column already zero below the diagonal, Q formed
// by applying the reflectors to the identity in reverse order, reduced or full shapes.
//
// How: operands are brought into contiguous compute-type work matrices by the backend's own
// cast / copy kernels (any strides, any storage dtype), ONE CTA per batch matrix then runs the
// factorization with the matrix in global memory (L2-resident at these sizes): right-looking
// rank-1 updates for cholesky and the substitution (parallel over the trailing block, coalesced
// along rows), one thread per column for the Householder applies -- the reference's unblocked
// operation order, so small matrices agree to rounding. Results go back through cast / copy.
// A device status word reports numeric failure; the host reads it (one 4-byte read-back) because
// the reference raises synchronously. Large single matrices want the blocked, GEMM-fed variant;
// this one is sized for the batched small/medium factorizations ML code issues.
#include <cooperative_groups.h>
#include "nxc_ops.cuh"

#define NXC_LA_THREADS 512
#include "nxc_linalg3.cuh"  // svd / eig / eigh kernel bodies (shared with the host emulation in tests/emu)

static const char NXC_LA_NOT_PD[] = "matrix is not positive definite";
static const char NXC_LA_SINGULAR[] = "triangular matrix is singular";
static const char NXC_LA_NOT_FLOAT[] = "linalg requires a float or complex dtype";
static const char NXC_LA_NOT_SQUARE[] = "matrix must be square";
static const char NXC_LA_SHAPE[] = "operand shapes are incompatible";

// ---- scalar algebra over the four compute types ------------------------------------------------
template <class T> struct LA;
template <> struct LA<float> {
  typedef float R;
  __device__ static float conj(float a) { return a; }
  __device__ static float norm2(float a) { return a * a; }
  __device__ static float real(float a) { return a; }
  __device__ static float imag(float) { return 0.f; }
  __device__ static float mk(float re, float) { return re; }
  __device__ static float add(float a, float b) { return a + b; }
  __device__ static float sub(float a, float b) { return a - b; }
  __device__ static float mul(float a, float b) { return a * b; }
  __device__ static float div(float a, float b) { return a / b; }
  __device__ static float divr(float a, float r) { return a / r; }
  __device__ static bool is_zero(float a) { return a == 0.f; }
  __device__ static float rsqrt_(float a) { return sqrtf(a); }
};
template <> struct LA<double> {
  typedef double R;
  __device__ static double conj(double a) { return a; }
  __device__ static double norm2(double a) { return a * a; }
  __device__ static double real(double a) { return a; }
  __device__ static double imag(double) { return 0.0; }
  __device__ static double mk(double re, double) { return re; }
  __device__ static double add(double a, double b) { return a + b; }
  __device__ static double sub(double a, double b) { return a - b; }
  __device__ static double mul(double a, double b) { return a * b; }
  __device__ static double div(double a, double b) { return a / b; }
  __device__ static double divr(double a, double r) { return a / r; }
  __device__ static bool is_zero(double a) { return a == 0.0; }
  __device__ static double rsqrt_(double a) { return sqrt(a); }
};
template <class Z, class R_> struct LAZ {
  typedef R_ R;
  __device__ static Z conj(Z a) { return zmk<Z>(a.re, -a.im); }
  __device__ static R norm2(Z a) { return a.re * a.re + a.im * a.im; }
  __device__ static R real(Z a) { return a.re; }
  __device__ static R imag(Z a) { return a.im; }
  __device__ static Z mk(R re, R im) { return zmk<Z>(re, im); }
  __device__ static Z add(Z a, Z b) { return zadd(a, b); }
  __device__ static Z sub(Z a, Z b) { return zsub(a, b); }
  __device__ static Z mul(Z a, Z b) { return zmk<Z>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
  __device__ static Z div(Z a, Z b) { return zdiv(a, b); }
  __device__ static Z divr(Z a, R r) { return zmk<Z>(a.re / r, a.im / r); }
  __device__ static bool is_zero(Z a) { return a.re == (R)0 && a.im == (R)0; }
};
template <> struct LA<cf32> : LAZ<cf32, float> { __device__ static float rsqrt_(float a) { return sqrtf(a); } };
template <> struct LA<cf64> : LAZ<cf64, double> { __device__ static double rsqrt_(double a) { return sqrt(a); } };

// CTA-wide sum of one real per thread (result broadcast to every thread)
template <class R>
__device__ R nxc_la_block_sum(R v, R *sm) {
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  R t = (R)0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sm[w];
  return t;
}

// ---- cholesky: A = L L^H, lower, right-looking, in place on an n x n row-major work matrix -----
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS) nxc_cholesky_kernel(T *work, int64_t n, int upper, int *status) {
  typedef LA<T> L;
  typedef typename L::R R;
  T *A = work + (int64_t)blockIdx.x * n * n;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int64_t j = 0; j < n; j++) {
    // pivot: every earlier column's update has already been subtracted from A[j][j]
    if (threadIdx.x == 0) {
      const R d = L::real(A[j * n + j]);
      if (!(d > (R)0)) { bad = 1; atomicExch(status, 1); }
      else A[j * n + j] = L::mk(L::rsqrt_(d), (R)0);
    }
    __syncthreads();
    if (bad) return;
    const R ljj = L::real(A[j * n + j]);
    for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[i * n + j] = L::divr(A[i * n + j], ljj);
    __syncthreads();
    // trailing lower triangle: A[i][c] -= L[i][j] * conj(L[c][j]),  j < c <= i
    const int64_t t = n - 1 - j;
    for (int64_t e = threadIdx.x; e < t * t; e += blockDim.x) {
      const int64_t i = j + 1 + e / t, c = j + 1 + e % t;
      if (c <= i) A[i * n + c] = L::sub(A[i * n + c], L::mul(A[i * n + j], L::conj(A[c * n + j])));
    }
    __syncthreads();
  }
  // the other triangle: zeros for ~lower; ~upper returns L^H and zeros below
  for (int64_t e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int64_t i = e / n, c = e % n;
    if (c > i) A[i * n + c] = upper ? L::conj(A[c * n + i]) : L::mk((R)0, (R)0);
  }
  if (upper) {
    __syncthreads();
    for (int64_t e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int64_t i = e / n, c = e % n;
      if (c < i) A[i * n + c] = L::mk((R)0, (R)0);
    }
  }
}

// ---- blocked cholesky for large real matrices ----------------------------------------------------
// One CTA per matrix is the right shape for the batched small systems the reference's callers
// issue; a single 512 x 512 matrix took 25 ms that way.  Large real matrices go right-looking in
// panels of NXC_CH_NB columns: (1) the diagonal block is factored in shared memory by one CTA per
// matrix, (2) the panel below it is solved row by row against that block, one thread per row,
// (3) the trailing update A22 -= L21 L21^T is a product through nxc_matmul plus one strided
// subtract -- all on views of the same work matrix, no block is copied out.  Same contract as the
// kernel above: lower triangle read, status 1 when a pivot is not > 0.
#define NXC_CH_NB 64
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS)
nxc_chol_diag_kernel(T *work, int64_t n, int64_t j0, int nb, int *status) {
  __shared__ T S[NXC_CH_NB][NXC_CH_NB + 1];
  __shared__ int bad;
  T *A = work + (int64_t)blockIdx.x * n * n + j0 * n + j0;
  if (threadIdx.x == 0) bad = 0;
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) S[e / nb][e % nb] = A[(int64_t)(e / nb) * n + e % nb];
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    if (threadIdx.x == 0) {
      const T d = S[j][j];
      if (!(d > (T)0)) { bad = 1; atomicExch(status, 1); }
      else S[j][j] = sqrt(d);
    }
    __syncthreads();
    if (bad) return;
    const T ljj = S[j][j];
    for (int i = j + 1 + threadIdx.x; i < nb; i += blockDim.x) S[i][j] = S[i][j] / ljj;
    __syncthreads();
    const int t = nb - 1 - j;
    for (int e = threadIdx.x; e < t * t; e += blockDim.x) {
      const int i = j + 1 + e / t, c = j + 1 + e % t;
      if (c <= i) S[i][c] -= S[i][j] * S[c][j];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x)
    if (e % nb <= e / nb) A[(int64_t)(e / nb) * n + e % nb] = S[e / nb][e % nb];
}

// rows below a full diagonal block: x L11^T = a, forward over the block's columns
template <class T>
__global__ void __launch_bounds__(128) nxc_chol_panel_kernel(T *work, int64_t n, int64_t j0) {
  __shared__ T S[NXC_CH_NB][NXC_CH_NB + 1];
  T *M = work + (int64_t)blockIdx.y * n * n;
  for (int e = threadIdx.x; e < NXC_CH_NB * NXC_CH_NB; e += blockDim.x)
    S[e / NXC_CH_NB][e % NXC_CH_NB] = M[(j0 + e / NXC_CH_NB) * n + j0 + e % NXC_CH_NB];
  __syncthreads();
  const int64_t row = j0 + NXC_CH_NB + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  T *a = M + row * n + j0;
  T x[NXC_CH_NB];
#pragma unroll
  for (int c = 0; c < NXC_CH_NB; c++) {
    T v = a[c];
#pragma unroll
    for (int k = 0; k < c; k++) v -= x[k] * S[c][k];
    x[c] = v / S[c][c];
  }
#pragma unroll
  for (int c = 0; c < NXC_CH_NB; c++) a[c] = x[c];
}

// the other triangle: zeros for ~lower; ~upper moves L to L^T and zeroes below
template <class T>
__global__ void __launch_bounds__(256) nxc_chol_finish_kernel(T *work, int64_t n, int upper) {
  T *M = work + (int64_t)blockIdx.y * n * n;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const int64_t i = e / n, c = e % n;
  if (c >= i) return;
  if (upper) { M[c * n + i] = M[i * n + c]; M[i * n + c] = (T)0; }
  else M[c * n + i] = (T)0;
}

// ---- triangular solve: op(A) X = B in place on X (n x nrhs), A n x n ----------------------------
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS)
nxc_trsm_kernel(const T *aw, T *xw, int64_t n, int64_t nrhs, int upper, int transpose, int unit, int *status) {
  typedef LA<T> L;
  const T *A = aw + (int64_t)blockIdx.x * n * n;
  T *X = xw + (int64_t)blockIdx.x * n * nrhs;
  const bool forward = (upper != 0) == (transpose != 0);
  for (int64_t ii = 0; ii < n; ii++) {
    const int64_t i = forward ? ii : n - 1 - ii;
    const T diag = transpose ? L::conj(A[i * n + i]) : A[i * n + i];
    if (!unit && L::is_zero(diag)) {  // uniform across the CTA
      if (threadIdx.x == 0) atomicExch(status, 2);
      return;
    }
    if (!unit) {
      for (int64_t j = threadIdx.x; j < nrhs; j += blockDim.x) X[i * nrhs + j] = L::div(X[i * nrhs + j], diag);
      __syncthreads();
    }
    // rows not yet solved lose their coupling to row i
    const int64_t rest = n - 1 - ii;
    for (int64_t e = threadIdx.x; e < rest * nrhs; e += blockDim.x) {
      const int64_t q = e / nrhs, j = e - q * nrhs;
      const int64_t r = forward ? i + 1 + q : i - 1 - q;
      const T c = transpose ? L::conj(A[i * n + r]) : A[r * n + i];
      X[r * nrhs + j] = L::sub(X[r * nrhs + j], L::mul(c, X[i * nrhs + j]));
    }
    __syncthreads();
  }
}

// ---- blocked triangular solve for large real systems ---------------------------------------------
// Same panel shape as the blocked cholesky: M = op(A) is walked in NXC_CH_NB-row blocks in
// substitution order; a block of X is solved against the diagonal block of M held in shared memory,
// one thread per right-hand-side column (coalesced along X's rows), and the rows not yet solved lose
// their coupling to the block through one product: X[rest] -= M[rest, blk] X[blk].  M is addressed
// through its row / column strides, so `transpose` is a stride swap (real dtypes only).
template <class T>
__global__ void __launch_bounds__(128)
nxc_trsm_diag_kernel(const T *aw, T *xw, int64_t n, int64_t nrhs, int64_t mrs, int64_t mcs, int64_t j0, int nb,
                     int forward, int unit, int *status) {
  __shared__ T S[NXC_CH_NB][NXC_CH_NB + 1];
  const T *M = aw + (int64_t)blockIdx.y * n * n;
  for (int e = threadIdx.x; e < NXC_CH_NB * NXC_CH_NB; e += blockDim.x) {
    const int i = e / NXC_CH_NB, c = e % NXC_CH_NB;
    // a ragged block is padded with the identity: the padded unknowns stay zero
    S[i][c] = (i < nb && c < nb) ? M[(j0 + i) * mrs + (j0 + c) * mcs] : (T)(i == c ? 1 : 0);
  }
  __syncthreads();
  if (!unit && blockIdx.x == 0 && threadIdx.x < nb && S[threadIdx.x][threadIdx.x] == (T)0) atomicExch(status, 2);
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nrhs) return;
  T *X = xw + (int64_t)blockIdx.y * n * nrhs + j0 * nrhs + col;
  T x[NXC_CH_NB];
#pragma unroll
  for (int c = 0; c < NXC_CH_NB; c++) x[c] = c < nb ? X[(int64_t)c * nrhs] : (T)0;
  if (forward) {
#pragma unroll
    for (int c = 0; c < NXC_CH_NB; c++) {
      T v = x[c];
#pragma unroll
      for (int k = 0; k < c; k++) v -= S[c][k] * x[k];
      x[c] = unit ? v : v / S[c][c];
    }
  } else {
#pragma unroll
    for (int c = NXC_CH_NB - 1; c >= 0; c--) {
      T v = x[c];
#pragma unroll
      for (int k = c + 1; k < NXC_CH_NB; k++) v -= S[c][k] * x[k];
      x[c] = unit ? v : v / S[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < NXC_CH_NB; c++)
    if (c < nb) X[(int64_t)c * nrhs] = x[c];
}

// ---- Householder QR on an m x n work matrix; Q (m x nq) formed from the reflectors ----------------
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS)
nxc_qr_kernel(T *work, T *qbuf, T *taubuf, int64_t m, int64_t n, int64_t nq) {
  typedef LA<T> L;
  typedef typename L::R R;
  T *A = work + (int64_t)blockIdx.x * m * n;
  T *Q = qbuf + (int64_t)blockIdx.x * m * nq;
  const int64_t k = m < n ? m : n;
  T *tau = taubuf + (int64_t)blockIdx.x * (k > 0 ? k : 1);
  __shared__ R red[NXC_LA_THREADS / 32];
  __shared__ T s_tau;
  for (int64_t j = 0; j < k; j++) {
    R part = (R)0;
    for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) part += L::norm2(A[i * n + j]);
    const R xnorm2 = nxc_la_block_sum<R>(part, red);
    const T alpha = A[j * n + j];
    const R alphr = L::real(alpha);
    if (xnorm2 == (R)0) {  // uniform
      if (threadIdx.x == 0) tau[j] = L::mk((R)0, (R)0);
      __syncthreads();
      continue;
    }
    const R anorm = L::rsqrt_(L::norm2(alpha) + xnorm2);
    const R beta = alphr >= (R)0 ? -anorm : anorm;
    const T tj = L::mk((beta - alphr) / beta, -L::imag(alpha) / beta);
    const T scal = L::sub(alpha, L::mk(beta, (R)0));
    __syncthreads();  // everyone has read alpha
    for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) A[i * n + j] = L::div(A[i * n + j], scal);
    if (threadIdx.x == 0) { A[j * n + j] = L::mk(beta, (R)0); tau[j] = tj; s_tau = tj; }
    __syncthreads();
    // apply H_j^H to the columns right of j: one thread per column, rows in order
    const T tauc = L::conj(s_tau);
    for (int64_t c = j + 1 + threadIdx.x; c < n; c += blockDim.x) {
      T w = A[j * n + c];
      for (int64_t i = j + 1; i < m; i++) w = L::add(w, L::mul(L::conj(A[i * n + j]), A[i * n + c]));
      w = L::mul(tauc, w);
      A[j * n + c] = L::sub(A[j * n + c], w);
      for (int64_t i = j + 1; i < m; i++) A[i * n + c] = L::sub(A[i * n + c], L::mul(w, A[i * n + j]));
    }
    __syncthreads();
  }
  // Q = H_0 ... H_{k-1} applied to the identity, last reflector first
  for (int64_t e = threadIdx.x; e < m * nq; e += blockDim.x) {
    const int64_t i = e / nq, c = e - i * nq;
    Q[e] = i == c ? L::mk((R)1, (R)0) : L::mk((R)0, (R)0);
  }
  __syncthreads();
  for (int64_t jj = 0; jj < k; jj++) {
    const int64_t j = k - 1 - jj;
    const T tj = tau[j];
    for (int64_t c = threadIdx.x; c < nq; c += blockDim.x) {
      T w = Q[j * nq + c];
      for (int64_t i = j + 1; i < m; i++) w = L::add(w, L::mul(L::conj(A[i * n + j]), Q[i * nq + c]));
      w = L::mul(tj, w);
      Q[j * nq + c] = L::sub(Q[j * nq + c], w);
      for (int64_t i = j + 1; i < m; i++) Q[i * nq + c] = L::sub(Q[i * nq + c], L::mul(w, A[i * n + j]));
    }
    __syncthreads();
  }
  // R: the upper trapezoid; the reflector tails below the diagonal become zeros
  for (int64_t e = threadIdx.x; e < m * n; e += blockDim.x) {
    const int64_t i = e / n, c = e - i * n;
    if (c < i) A[e] = L::mk((R)0, (R)0);
  }
}


// ---- blocked Householder QR for large real matrices (compact WY) ---------------------------------
// The one-CTA kernel above applies each reflector to the whole matrix with one thread per column:
// 72 ms for one 512 x 512 matrix.  Large real matrices are factored in panels of NXC_QR_NB columns:
// one CTA per matrix factors the panel with the same reflector arithmetic (LAPACK pivot sign,
// tau = 0 for an already-reduced column) -- threads laid out as 32 panel columns x 16 row lanes, so
// a warp reads one 128-byte panel row -- and accumulates the triangular factor T of the block
// reflector H_0 ... H_{nb-1} = I - V T V^T in the same pass that applies H_j (the dot products
// V[:, c]^T v_j for c < j are the T column, for c > j the update).  The panel's explicit V (unit
// lower trapezoid) goes to its own buffer, and everything right of the panel -- and all of Q,
// panels in reverse -- is three products through nxc_matmul: X -= V (T' (V^T X)).
#define NXC_QR_NB 32
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS)
nxc_qr_panel_kernel(T *work, T *vall, T *tall, int64_t m, int64_t n, int64_t k, int64_t j0, int nb, int64_t npanels) {
  constexpr int LANES = NXC_LA_THREADS / NXC_QR_NB;
  __shared__ T sT[NXC_QR_NB][NXC_QR_NB + 1];
  __shared__ T sw[LANES][NXC_QR_NB];
  __shared__ T wv[NXC_QR_NB];
  __shared__ T red[NXC_LA_THREADS / 32];
  T *A = work + (int64_t)blockIdx.x * m * n;
  T *V = vall + (int64_t)blockIdx.x * m * k;
  T *Tm = tall + ((int64_t)blockIdx.x * npanels + j0 / NXC_QR_NB) * NXC_QR_NB * NXC_QR_NB;
  const int c = threadIdx.x % NXC_QR_NB, r = threadIdx.x / NXC_QR_NB;
  for (int e = threadIdx.x; e < NXC_QR_NB * NXC_QR_NB; e += blockDim.x) sT[e / NXC_QR_NB][e % NXC_QR_NB] = (T)0;
  __syncthreads();
  for (int jj = 0; jj < nb; jj++) {
    const int64_t j = j0 + jj;
    T part = (T)0;
    for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) part += A[i * n + j] * A[i * n + j];
    const T xnorm2 = nxc_la_block_sum<T>(part, red);
    const T alpha = A[j * n + j];
    T tau = (T)0;
    if (xnorm2 != (T)0) {  // uniform
      const T anorm = sqrt(alpha * alpha + xnorm2);
      const T beta = alpha >= (T)0 ? -anorm : anorm;
      tau = (beta - alpha) / beta;
      const T scal = alpha - beta;
      __syncthreads();  // everyone has read alpha
      for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) A[i * n + j] = A[i * n + j] / scal;
      if (threadIdx.x == 0) A[j * n + j] = beta;
    }
    __syncthreads();
    // w_c = V[j][c] + sum_{i > j} v_i A[i][c] over the panel's columns
    T acc = (T)0;
    if (c < nb && c != jj)
      for (int64_t i = j + 1 + r; i < m; i += LANES) acc += A[i * n + j] * A[i * n + j0 + c];
    sw[r][c] = acc;
    __syncthreads();
    if (threadIdx.x < nb) {
      T w = A[j * n + j0 + threadIdx.x];
      for (int q = 0; q < LANES; q++) w += sw[q][threadIdx.x];
      wv[threadIdx.x] = w;
    }
    __syncthreads();
    // T[0:jj, jj] = -tau T[0:jj, 0:jj] (V[:, 0:jj]^T v_jj),  T[jj][jj] = tau
    if (threadIdx.x < jj) {
      T sacc = (T)0;
      for (int b = threadIdx.x; b < jj; b++) sacc += sT[threadIdx.x][b] * wv[b];
      sT[threadIdx.x][jj] = -tau * sacc;
    } else if (threadIdx.x == jj) {
      sT[jj][jj] = tau;
    }
    // H_jj on the panel's remaining columns
    if (tau != (T)0 && c > jj && c < nb) {
      const T wc = tau * wv[c];
      if (r == 0) A[j * n + j0 + c] -= wc;
      for (int64_t i = j + 1 + r; i < m; i += LANES) A[i * n + j0 + c] -= wc * A[i * n + j];
    }
    __syncthreads();
  }
  // the explicit unit lower trapezoid and the triangular factor
  for (int64_t i = j0 + r; i < m; i += LANES)
    if (c < nb) V[i * k + j0 + c] = i < j0 + c ? (T)0 : (i == j0 + c ? (T)1 : A[i * n + j0 + c]);
  for (int e = threadIdx.x; e < NXC_QR_NB * NXC_QR_NB; e += blockDim.x) Tm[e] = sT[e / NXC_QR_NB][e % NXC_QR_NB];
}

// The same panel on a thread-block CLUSTER: the single CTA above walks an m x 32 panel through L2
// three times per column (1.0 ms per panel at m = 4096).  Eight CTAs of a cluster each keep m / 8
// panel rows in shared memory for the whole panel; the two reductions a column needs (the norm of
// its tail, then the 32 dot products against v_j) are finished by every CTA writing its partial into
// every peer's shared memory (DSMEM) and one cluster barrier -- each CTA then adds the eight partials
// in rank order, so all of them derive the same tau / T and no value is ever broadcast back.
#define NXC_QR_CS 8
#define NXC_QR_PANEL_SMEM_MAX (200u * 1024u)
template <class T>
__global__ void __launch_bounds__(NXC_LA_THREADS)
nxc_qr_panel_cluster_kernel(T *work, T *vall, T *tall, int64_t m, int64_t n, int64_t k, int64_t j0, int nb, int64_t npanels,
                            int rpc) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int LANES = NXC_LA_THREADS / NXC_QR_NB, NB = NXC_QR_NB;
  extern __shared__ __align__(16) unsigned char nxc_qr_smem[];
  T *P = (T *)nxc_qr_smem;  // [rpc][NB]: this CTA's rows of the panel
  __shared__ T sT[NB][NB + 1];
  __shared__ T sw[LANES][NB];
  __shared__ T wv[NB];
  __shared__ T red[NXC_LA_THREADS / 32];
  __shared__ T xn[NXC_QR_CS][2];   // per rank: partial tail norm; [owner][1] = the pivot
  __shared__ T wpart[NXC_QR_CS][NB];
  const int q = (int)cluster.block_rank();
  const int64_t b = blockIdx.x / NXC_QR_CS;
  T *A = work + b * m * n;
  T *V = vall + b * m * k;
  const int c = threadIdx.x % NB, r = threadIdx.x / NB;
  const int64_t r_lo = j0 + (int64_t)q * rpc;
  const int nloc = (int)(r_lo >= m ? 0 : (m - r_lo < rpc ? m - r_lo : rpc));
  for (int e = threadIdx.x; e < nloc * NB; e += blockDim.x)
    P[e] = (e % NB) < nb ? A[(r_lo + e / NB) * n + j0 + e % NB] : (T)0;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) sT[e / NB][e % NB] = (T)0;
  __syncthreads();
  cluster.sync();  // every peer is resident before its shared memory is written
  for (int jj = 0; jj < nb; jj++) {
    const int64_t j = j0 + jj;
    const int owner = (int)((j - j0) / rpc);  // the rank holding row j
    const int jl = (int)(j - r_lo);           // row j's local index there
    const int lo = (int)(j + 1 - r_lo < 0 ? 0 : (j + 1 - r_lo > nloc ? nloc : j + 1 - r_lo));
    T part = (T)0;
    for (int li = lo + threadIdx.x; li < nloc; li += blockDim.x) part += P[li * NB + jj] * P[li * NB + jj];
    const T mine = nxc_la_block_sum<T>(part, red);
    if (threadIdx.x < NXC_QR_CS) {
      T *peer = cluster.map_shared_rank(&xn[0][0], threadIdx.x);
      peer[q * 2] = mine;
      if (q == owner) peer[q * 2 + 1] = P[jl * NB + jj];
    }
    cluster.sync();
    T xnorm2 = (T)0;
    for (int t = 0; t < NXC_QR_CS; t++) xnorm2 += xn[t][0];
    const T alpha = xn[owner][1];
    T tau = (T)0;
    if (xnorm2 != (T)0) {  // uniform across the cluster
      const T anorm = sqrt(alpha * alpha + xnorm2);
      const T beta = alpha >= (T)0 ? -anorm : anorm;
      tau = (beta - alpha) / beta;
      const T scal = alpha - beta;
      for (int li = lo + threadIdx.x; li < nloc; li += blockDim.x) P[li * NB + jj] = P[li * NB + jj] / scal;
      if (q == owner && threadIdx.x == 0) P[jl * NB + jj] = beta;
    }
    __syncthreads();
    T acc = (T)0;
    if (c < nb && c != jj)
      for (int li = lo + r; li < nloc; li += LANES) acc += P[li * NB + jj] * P[li * NB + c];
    sw[r][c] = acc;
    __syncthreads();
    if (threadIdx.x < NB) {
      T w = (q == owner && threadIdx.x < nb) ? P[jl * NB + threadIdx.x] : (T)0;
      for (int t = 0; t < LANES; t++) w += sw[t][threadIdx.x];
      for (int t = 0; t < NXC_QR_CS; t++) cluster.map_shared_rank(&wpart[0][0], t)[q * NB + threadIdx.x] = w;
    }
    cluster.sync();
    if (threadIdx.x < NB) {
      T w = (T)0;
      for (int t = 0; t < NXC_QR_CS; t++) w += wpart[t][threadIdx.x];
      wv[threadIdx.x] = w;
    }
    __syncthreads();
    if (threadIdx.x < jj) {
      T sacc = (T)0;
      for (int bb = threadIdx.x; bb < jj; bb++) sacc += sT[threadIdx.x][bb] * wv[bb];
      sT[threadIdx.x][jj] = -tau * sacc;
    } else if (threadIdx.x == jj) {
      sT[jj][jj] = tau;
    }
    if (tau != (T)0 && c > jj && c < nb) {
      const T wc = tau * wv[c];
      if (q == owner && r == 0) P[jl * NB + c] -= wc;
      for (int li = lo + r; li < nloc; li += LANES) P[li * NB + c] -= wc * P[li * NB + jj];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nloc * NB; e += blockDim.x) {
    const int64_t i = r_lo + e / NB;
    const int cc = e % NB;
    if (cc < nb) {
      A[i * n + j0 + cc] = P[e];
      V[i * k + j0 + cc] = i < j0 + cc ? (T)0 : (i == j0 + cc ? (T)1 : P[e]);
    }
  }
  if (q == 0) {
    T *Tm = tall + (b * npanels + j0 / NB) * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Tm[e] = sT[e / NB][e % NB];
  }
}

// Q starts as the identity's first nq columns; R keeps the upper trapezoid
template <class T>
__global__ void __launch_bounds__(256) nxc_qr_eye_kernel(T *q, int64_t m, int64_t nq) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m * nq) q[(int64_t)blockIdx.y * m * nq + e] = e / nq == e % nq ? (T)1 : (T)0;
}
template <class T>
__global__ void __launch_bounds__(256) nxc_qr_triu_kernel(T *w, int64_t m, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m * n && e % n < e / n) w[(int64_t)blockIdx.y * m * n + e] = (T)0;
}

template <class T> struct La3Map;
template <> struct La3Map<float> { typedef float E; };
template <> struct La3Map<double> { typedef double E; };
template <> struct La3Map<cf32> { typedef La3C32 E; };
template <> struct La3Map<cf64> { typedef La3C64 E; };

__device__ __forceinline__ La3Thr nxc_la3_thr() {
  La3Thr t;
  t.tid = (int)threadIdx.x; t.nt = (int)blockDim.x;
  t.lane = (int)(threadIdx.x & 31); t.lanes = 32;
  t.warp = (int)(threadIdx.x >> 5); t.nwarps = (int)(blockDim.x >> 5);
  t.wide = 0;
  return t;
}

// The Jacobi kernels (eigh, svd) give a WARP a column pair and need a barrier per round-robin
// step; one CTA has 16 warps for the n / 2 pairs of a step.  Launched as a thread-block cluster
// (8 CTAs, 16 where the device schedules it) the team is 128 / 256 warps: all pairs of a 512-wide
// matrix rotate at once, and the step barrier is the cluster's hardware barrier.  The working
// matrices already live in global memory (L2); only the reduction scratch and the rotation counter
// move from shared to global memory.
__device__ __forceinline__ La3Thr nxc_la3_team(unsigned *rank, unsigned *size) {
  unsigned r, n;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
  La3Thr t;
  t.tid = (int)(r * blockDim.x + threadIdx.x); t.nt = (int)(n * blockDim.x);
  t.lane = (int)(threadIdx.x & 31); t.lanes = 32;
  t.warp = (int)(r * (blockDim.x >> 5) + (threadIdx.x >> 5)); t.nwarps = (int)(n * (blockDim.x >> 5));
  t.wide = n > 1;
  *rank = r; *size = n;
  return t;
}
// per-matrix global scratch of a cluster team: NXC_LA_TEAM_RED doubles + 2 ints
#define NXC_LA_TEAM_RED 256
#define NXC_LA_TEAM_BYTES (NXC_LA_TEAM_RED * 8 + 64)

// cluster width for `nbatch` matrices with `pairs` column pairs per step: 1 = plain launch
template <class K>
static int nxc_la_team_size(K kernel, int64_t nbatch, int64_t pairs) {
  int cs = 1;
  const char *env = getenv("NX_CUDA_LA_CLUSTER");
  if (env) {
    cs = atoi(env);
    if (cs != 2 && cs != 4 && cs != 8 && cs != 16) cs = 1;
  } else if (pairs >= 64) {
    // widen while the batch leaves SMs free and a step still has a pair for every warp
    while (cs < 16 && nbatch * cs * 2 <= 148 && pairs > (int64_t)cs * (NXC_LA_THREADS / 32)) cs *= 2;
  }
  if (cs == 16) {
    // more than 8 CTAs per cluster is opt-in, and only where the device can place one
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(NXC_LA_THREADS); cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      cs = 8;
    }
  }
  return cs;
}

template <class K, class A>
static cudaError_t nxc_la_team_launch(nxc_ctx *ctx, K kernel, int cs, int64_t nbatch, const A &args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(nbatch * cs));
  cfg.blockDim = dim3(NXC_LA_THREADS);
  cfg.stream = ctx->stream;
  cfg.attrs = attr;
  cfg.numAttrs = cs > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args);
}

struct NxcEighArgs {
  const void *a;
  void *gt, *vt, *vo;
  double *w, *sg;
  int *rk;
  int64_t n;
  int vectors;
  int *status;
  char *team;  // NXC_LA_TEAM_BYTES per matrix when launched as clusters
};

template <class E>
__global__ void __launch_bounds__(NXC_LA_THREADS) nxc_eigh_kernel(const NxcEighArgs a) {
  __shared__ double sred[NXC_LA_THREADS / 32];
  __shared__ int sflags[2];
  unsigned rank, size;
  const La3Thr t = nxc_la3_team(&rank, &size);
  const int64_t b = blockIdx.x / size, nn = a.n * a.n;
  double *red = t.wide ? (double *)(a.team + b * NXC_LA_TEAM_BYTES) : sred;
  int *flags = t.wide ? (int *)(a.team + b * NXC_LA_TEAM_BYTES + NXC_LA_TEAM_RED * 8) : sflags;
  la3_eigh_body<E>(t, (const E *)a.a + b * nn, (E *)a.gt + b * nn, (E *)a.vt + b * nn, (E *)a.vo + b * nn,
                   a.w + b * a.n, a.sg + b * a.n, a.rk + b * a.n, red, flags, a.n, a.vectors, 60, a.status);
}

// ---- host side -----------------------------------------------------------------------------------
static int nxc_la_compute_dtype(int dt) {
  switch (dt) {
    case NXC_F16: case NXC_BF16: case NXC_F8E4M3: case NXC_F8E5M2: case NXC_F32: return NXC_F32;
    case NXC_F64: return NXC_F64;
    case NXC_C32: return NXC_C32;
    case NXC_C64: return NXC_C64;
    default: return -1;
  }
}

// contiguous [batch..., rows, cols] descriptor over a fresh work buffer
static nxc_status nxc_la_work(nxc_ctx *ctx, const nxc_tensor *like, int cdt, int64_t rows, int64_t cols, nxc_tensor *w,
                              void **buf) {
  *w = *like;
  w->dtype = cdt;
  w->offset = 0;
  w->shape[w->ndim - 2] = rows;
  w->shape[w->ndim - 1] = cols;
  int64_t st = 1;
  for (int d = w->ndim - 1; d >= 0; d--) { w->strides[d] = st; st *= w->shape[d]; }
  nxc_status s = nxc_alloc(ctx, (size_t)(st > 0 ? st : 1) * (size_t)nxc_elem_size(cdt), buf);
  w->data = *buf;
  return s;
}

// storage -> work (cast when the dtypes differ, else copy); work -> storage likewise
static nxc_status nxc_la_move(nxc_ctx *ctx, const nxc_tensor *dst, const nxc_tensor *src) {
  return dst->dtype == src->dtype ? nxc_copy(ctx, dst, src) : nxc_cast(ctx, dst, src);
}

static nxc_status nxc_la_status(nxc_ctx *ctx, int *dev_status) {
  int h = 0;
  NXC_CUDA_TRY(ctx, cudaMemcpyAsync(&h, dev_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h == 1) return NXC_LA_NOT_PD;
  if (h == 2) return NXC_LA_SINGULAR;
  return NXC_OK;
}

static nxc_status nxc_la_fail(nxc_ctx *ctx, nxc_status s) {
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}

#define NXC_LA_DISPATCH(cdt, ...)                                         \
  switch (cdt) {                                                          \
    case NXC_F32: { typedef float T; __VA_ARGS__ } break;                 \
    case NXC_F64: { typedef double T; __VA_ARGS__ } break;                \
    case NXC_C32: { typedef cf32 T; __VA_ARGS__ } break;                  \
    default: { typedef cf64 T; __VA_ARGS__ } break;                       \
  }

// blocked factorisation of the contiguous work matrices w [batch..., n, n] (real dtypes)
template <class T>
static nxc_status nxc_cholesky_blocked(nxc_ctx *ctx, const nxc_tensor *w, int64_t n, int64_t nbatch, int upper, int *st) {
  nxc_status s = NXC_OK;
  const int nd = w->ndim;
  for (int64_t j0 = 0; j0 < n && !s; j0 += NXC_CH_NB) {
    const int nb = (int)(n - j0 < NXC_CH_NB ? n - j0 : NXC_CH_NB);
    nxc_chol_diag_kernel<T><<<(unsigned)nbatch, NXC_LA_THREADS, 0, ctx->stream>>>((T *)w->data, n, j0, nb, st);
    ctx->launches++;
    const int64_t m = n - j0 - nb;
    if (m <= 0) break;
    nxc_chol_panel_kernel<T><<<dim3((unsigned)((m + 127) / 128), (unsigned)nbatch), 128, 0, ctx->stream>>>((T *)w->data, n, j0);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) return nxc_cuda_fail(ctx, cudaGetLastError(), "cholesky panel");
    // A22 -= L21 L21^T on views of the work matrix
    nxc_tensor l21 = *w, l21t = *w, a22 = *w, prod;
    void *pbuf = NULL;
    l21.offset = (j0 + nb) * n + j0;  l21.shape[nd - 2] = m;  l21.shape[nd - 1] = nb;
    l21t = l21;  l21t.shape[nd - 2] = nb;  l21t.shape[nd - 1] = m;  l21t.strides[nd - 2] = 1;  l21t.strides[nd - 1] = n;
    a22.offset = (j0 + nb) * n + (j0 + nb);  a22.shape[nd - 2] = m;  a22.shape[nd - 1] = m;
    if ((s = nxc_la_work(ctx, w, w->dtype, m, m, &prod, &pbuf))) return s;
    s = nxc_matmul(ctx, &prod, &l21, &l21t);
    if (!s) s = nxc_map2(ctx, NXC_SUB, &a22, &a22, &prod);
    nxc_free(ctx, pbuf);
  }
  if (s) return s;
  nxc_chol_finish_kernel<T><<<dim3((unsigned)((n * n + 255) / 256), (unsigned)nbatch), 256, 0, ctx->stream>>>((T *)w->data, n, upper);
  ctx->launches++;
  if (cudaPeekAtLastError() != cudaSuccess) return nxc_cuda_fail(ctx, cudaGetLastError(), "cholesky");
  return NXC_OK;
}

// where the panel form wins: measured on B200 (profiles/linalg_blocked_r02.json)
static bool nxc_cholesky_use_blocked(int cdt, int64_t n, int64_t nbatch) {
  if (cdt != NXC_F32 && cdt != NXC_F64) return false;  // complex keeps the one-CTA kernel (no conj-transposed product)
  if (getenv("NX_CUDA_CHOLESKY_BLOCKED")) return atoi(getenv("NX_CUDA_CHOLESKY_BLOCKED")) != 0 && n > NXC_CH_NB;
  (void)nbatch;  // the panel form also wins batched: 16 x 256^2 0.40 ms against 2.8 ms
  return n >= 128;
}

extern "C" nxc_status nxc_cholesky(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int upper) {
  NXC_TRACE(ctx, "nxc_cholesky");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_cholesky (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) return nxc_la_fail(ctx, s);
  if (in->ndim < 2 || out->ndim != in->ndim) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t n = in->shape[in->ndim - 1];
  if (in->shape[in->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_NOT_SQUARE);
  if (out->shape[out->ndim - 1] != n || out->shape[out->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int cdt = nxc_la_compute_dtype(in->dtype);
  if (cdt < 0) return nxc_la_fail(ctx, NXC_LA_NOT_FLOAT);
  int64_t nbatch = 1;
  for (int i = 0; i < in->ndim - 2; i++) {
    if (out->shape[i] != in->shape[i]) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= in->shape[i];
  }
  if (n == 0 || nbatch == 0) return NXC_OK;
  nxc_tensor w;
  void *buf = NULL;
  int *st = NULL;
  if ((s = nxc_la_work(ctx, in, cdt, n, n, &w, &buf))) return nxc_la_fail(ctx, s);
  s = nxc_alloc(ctx, sizeof(int), (void **)&st);
  if (!s) s = nxc_memset(ctx, st, 0, sizeof(int));
  if (!s) s = nxc_la_move(ctx, &w, in);
  if (!s && nxc_cholesky_use_blocked(cdt, n, nbatch)) {
    s = cdt == NXC_F32 ? nxc_cholesky_blocked<float>(ctx, &w, n, nbatch, upper, st)
                       : nxc_cholesky_blocked<double>(ctx, &w, n, nbatch, upper, st);
  } else if (!s) {
    NXC_LA_DISPATCH(cdt, { nxc_cholesky_kernel<T><<<(unsigned)nbatch, NXC_LA_THREADS, 0, ctx->stream>>>((T *)buf, n, upper, st); })
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "cholesky");
  }
  if (!s) s = nxc_la_status(ctx, st);
  // the reference leaves a failed matrix unwritten and still writes the others; a failed call's
  // output is unspecified either way, so nothing is written back on failure
  if (!s) s = nxc_la_move(ctx, out, &w);
  nxc_free(ctx, buf);
  if (st) nxc_free(ctx, st);
  return nxc_la_fail(ctx, s);
}

// blocked substitution on the contiguous work matrices aw [batch..., n, n], xw [batch..., n, nrhs]
template <class T>
static nxc_status nxc_trsm_blocked(nxc_ctx *ctx, const nxc_tensor *aw, const nxc_tensor *xw, int64_t n, int64_t nrhs,
                                   int64_t nbatch, int upper, int transpose, int unit, int *st) {
  const int nd = aw->ndim;
  const bool forward = (upper != 0) == (transpose != 0);
  const int64_t mrs = transpose ? 1 : n, mcs = transpose ? n : 1;
  const int64_t blocks = (n + NXC_CH_NB - 1) / NXC_CH_NB;
  for (int64_t bi = 0; bi < blocks; bi++) {
    const int64_t j0 = (forward ? bi : blocks - 1 - bi) * NXC_CH_NB;
    const int nb = (int)(n - j0 < NXC_CH_NB ? n - j0 : NXC_CH_NB);
    nxc_trsm_diag_kernel<T><<<dim3((unsigned)((nrhs + 127) / 128), (unsigned)nbatch), 128, 0, ctx->stream>>>(
        (const T *)aw->data, (T *)xw->data, n, nrhs, mrs, mcs, j0, nb, forward, unit, st);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) return nxc_cuda_fail(ctx, cudaGetLastError(), "triangular_solve block");
    const int64_t r0 = forward ? j0 + nb : 0, rest = forward ? n - j0 - nb : j0;
    if (rest <= 0) continue;
    nxc_tensor mv = *aw, xb = *xw, xr = *xw, prod;
    void *pbuf = NULL;
    mv.offset = r0 * mrs + j0 * mcs;  mv.shape[nd - 2] = rest;  mv.shape[nd - 1] = nb;
    mv.strides[nd - 2] = mrs;  mv.strides[nd - 1] = mcs;
    xb.offset = j0 * nrhs;  xb.shape[nd - 2] = nb;
    xr.offset = r0 * nrhs;  xr.shape[nd - 2] = rest;
    nxc_status s = nxc_la_work(ctx, xw, xw->dtype, rest, nrhs, &prod, &pbuf);
    if (!s) s = nxc_matmul(ctx, &prod, &mv, &xb);
    if (!s) s = nxc_map2(ctx, NXC_SUB, &xr, &xr, &prod);
    if (pbuf) nxc_free(ctx, pbuf);
    if (s) return s;
  }
  return NXC_OK;
}

static bool nxc_trsm_use_blocked(int cdt, int64_t n) {
  if (cdt != NXC_F32 && cdt != NXC_F64) return false;
  if (getenv("NX_CUDA_TRSM_BLOCKED")) return atoi(getenv("NX_CUDA_TRSM_BLOCKED")) != 0 && n > NXC_CH_NB;
  return n >= 128;
}

extern "C" nxc_status nxc_triangular_solve(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a, const nxc_tensor *b,
                                           int flags) {
  NXC_TRACE(ctx, "nxc_triangular_solve");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_triangular_solve (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(a)) || (s = nxc_check_tensor(b)) || (s = nxc_check_tensor(out))) return nxc_la_fail(ctx, s);
  if (a->ndim < 2 || b->ndim != a->ndim || out->ndim != b->ndim) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t n = a->shape[a->ndim - 1];
  if (a->shape[a->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_NOT_SQUARE);
  const int64_t nrhs = b->shape[b->ndim - 1];
  if (b->shape[b->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  if (out->shape[out->ndim - 2] != n || out->shape[out->ndim - 1] != nrhs) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int cdt = nxc_la_compute_dtype(a->dtype);
  if (cdt < 0) return nxc_la_fail(ctx, NXC_LA_NOT_FLOAT);
  if (b->dtype != a->dtype || out->dtype != a->dtype) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  int64_t nbatch = 1;
  for (int i = 0; i < a->ndim - 2; i++) {
    if (b->shape[i] != a->shape[i] || out->shape[i] != a->shape[i]) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= a->shape[i];
  }
  if (n == 0 || nrhs == 0 || nbatch == 0) return NXC_OK;
  nxc_tensor aw, xw;
  void *abuf = NULL, *xbuf = NULL;
  int *st = NULL;
  if ((s = nxc_la_work(ctx, a, cdt, n, n, &aw, &abuf))) return nxc_la_fail(ctx, s);
  s = nxc_la_work(ctx, b, cdt, n, nrhs, &xw, &xbuf);
  if (!s) s = nxc_alloc(ctx, sizeof(int), (void **)&st);
  if (!s) s = nxc_memset(ctx, st, 0, sizeof(int));
  if (!s) s = nxc_la_move(ctx, &aw, a);
  if (!s) s = nxc_la_move(ctx, &xw, b);
  if (!s && nxc_trsm_use_blocked(cdt, n)) {
    s = cdt == NXC_F32
            ? nxc_trsm_blocked<float>(ctx, &aw, &xw, n, nrhs, nbatch, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, st)
            : nxc_trsm_blocked<double>(ctx, &aw, &xw, n, nrhs, nbatch, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, st);
  } else if (!s) {
    NXC_LA_DISPATCH(cdt, {
      nxc_trsm_kernel<T><<<(unsigned)nbatch, NXC_LA_THREADS, 0, ctx->stream>>>((const T *)abuf, (T *)xbuf, n, nrhs, flags & 1,
                                                                               (flags >> 1) & 1, (flags >> 2) & 1, st);
    })
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "triangular_solve");
  }
  if (!s) s = nxc_la_status(ctx, st);
  if (!s) s = nxc_la_move(ctx, out, &xw);
  nxc_free(ctx, abuf);
  if (xbuf) nxc_free(ctx, xbuf);
  if (st) nxc_free(ctx, st);
  return nxc_la_fail(ctx, s);
}

// X[j0:, x0:] -= V_p op(T_p) (V_p^T X[j0:, x0:]) on views; X is a contiguous [batch..., xr, xc] work matrix
static nxc_status nxc_qr_apply_block(nxc_ctx *ctx, const nxc_tensor *x, int64_t xc, int64_t x0, const nxc_tensor *vall,
                                     const nxc_tensor *tall, int64_t m, int64_t k, int64_t p, int64_t j0, int nb, bool t_transposed) {
  const int nd = x->ndim;
  const int64_t mp = m - j0, cols = xc - x0;
  if (cols <= 0) return NXC_OK;
  nxc_tensor vp = *vall, vpt, xv = *x, tp = *tall, w1, w2, prod;
  void *b1 = NULL, *b2 = NULL, *b3 = NULL;
  vp.offset = j0 * k + j0;  vp.shape[nd - 2] = mp;  vp.shape[nd - 1] = nb;
  vpt = vp;  vpt.shape[nd - 2] = nb;  vpt.shape[nd - 1] = mp;  vpt.strides[nd - 2] = 1;  vpt.strides[nd - 1] = k;
  xv.offset = j0 * xc + x0;  xv.shape[nd - 2] = mp;  xv.shape[nd - 1] = cols;
  tp.offset = p * NXC_QR_NB * NXC_QR_NB;  tp.shape[nd - 2] = nb;  tp.shape[nd - 1] = nb;
  tp.strides[nd - 2] = t_transposed ? 1 : NXC_QR_NB;  tp.strides[nd - 1] = t_transposed ? NXC_QR_NB : 1;
  nxc_status s = nxc_la_work(ctx, x, x->dtype, nb, cols, &w1, &b1);
  if (!s) s = nxc_la_work(ctx, x, x->dtype, nb, cols, &w2, &b2);
  if (!s) s = nxc_la_work(ctx, x, x->dtype, mp, cols, &prod, &b3);
  if (!s) s = nxc_matmul(ctx, &w1, &vpt, &xv);
  if (!s) s = nxc_matmul(ctx, &w2, &tp, &w1);
  if (!s) s = nxc_matmul(ctx, &prod, &vp, &w2);
  if (!s) s = nxc_map2(ctx, NXC_SUB, &xv, &xv, &prod);
  if (b1) nxc_free(ctx, b1);
  if (b2) nxc_free(ctx, b2);
  if (b3) nxc_free(ctx, b3);
  return s;
}

// w [batch..., m, n] is factored in place (R in its upper trapezoid), qw [batch..., m, nq] receives Q
template <class T>
static nxc_status nxc_qr_blocked(nxc_ctx *ctx, const nxc_tensor *w, const nxc_tensor *qw, int64_t m, int64_t n, int64_t nq,
                                 int64_t nbatch) {
  const int64_t k = m < n ? m : n;
  const int64_t npanels = (k + NXC_QR_NB - 1) / NXC_QR_NB;
  nxc_tensor vall, tall;
  void *vbuf = NULL, *tbuf = NULL;
  nxc_status s = nxc_la_work(ctx, w, w->dtype, m, k, &vall, &vbuf);
  if (!s) s = nxc_la_work(ctx, w, w->dtype, npanels * NXC_QR_NB, NXC_QR_NB, &tall, &tbuf);
  // panels whose rows fit the cluster's shared memory (m - j0 <= 12800 f32 / 6400 f64) run there
  const bool use_cluster = !(getenv("NX_CUDA_QR_CLUSTER") && atoi(getenv("NX_CUDA_QR_CLUSTER")) == 0);
  const bool force_cluster = getenv("NX_CUDA_QR_CLUSTER") && atoi(getenv("NX_CUDA_QR_CLUSTER")) == 1;
  if (!s && use_cluster) {  // once per call: the attribute belongs to the device, a process may drive several
    cudaError_t e = cudaFuncSetAttribute(nxc_qr_panel_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)NXC_QR_PANEL_SMEM_MAX);
    if (e != cudaSuccess) s = nxc_cuda_fail(ctx, e, "qr panel attribute");
  }
  for (int64_t p = 0; p < npanels && !s; p++) {
    const int64_t j0 = p * NXC_QR_NB;
    const int nb = (int)(k - j0 < NXC_QR_NB ? k - j0 : NXC_QR_NB);
    const int64_t rpc = (m - j0 + NXC_QR_CS - 1) / NXC_QR_CS;
    const size_t panel_smem = (size_t)rpc * NXC_QR_NB * sizeof(T);
    // a cluster per matrix pays when the panel is tall and the batch leaves SMs idle (256 x 128^2: 0.86 ms with
    // one CTA per matrix, 3.5 ms with 2048 clustered CTAs)
    if (use_cluster && panel_smem <= NXC_QR_PANEL_SMEM_MAX && (force_cluster || (nbatch * NXC_QR_CS <= 148 && m - j0 >= 512))) {
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = NXC_QR_CS;  attr[0].val.clusterDim.y = 1;  attr[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)(nbatch * NXC_QR_CS));
      cfg.blockDim = dim3(NXC_LA_THREADS);
      cfg.dynamicSmemBytes = panel_smem;
      cfg.stream = ctx->stream;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, nxc_qr_panel_cluster_kernel<T>, (T *)w->data, (T *)vbuf, (T *)tbuf, m, n, k, j0, nb,
                                         npanels, (int)rpc);
      if (e != cudaSuccess) s = nxc_cuda_fail(ctx, e, "qr panel (cluster)");
    } else {
      nxc_qr_panel_kernel<T><<<(unsigned)nbatch, NXC_LA_THREADS, 0, ctx->stream>>>((T *)w->data, (T *)vbuf, (T *)tbuf, m, n, k, j0,
                                                                                   nb, npanels);
    }
    ctx->launches++;
    if (!s && cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "qr panel");
    // H_{nb-1} ... H_0 = (I - V T V^T)^T on the columns right of the panel
    if (!s) s = nxc_qr_apply_block(ctx, w, n, j0 + nb, &vall, &tall, m, k, p, j0, nb, true);
  }
  if (!s && nq > 0) {
    nxc_qr_eye_kernel<T><<<dim3((unsigned)((m * nq + 255) / 256), (unsigned)nbatch), 256, 0, ctx->stream>>>((T *)qw->data, m, nq);
    ctx->launches++;
    // Q = H_0 ... H_{k-1} I: block reflectors last to first; rows and columns before j0 are still the identity's
    for (int64_t p = npanels - 1; p >= 0 && !s; p--) {
      const int64_t j0 = p * NXC_QR_NB;
      const int nb = (int)(k - j0 < NXC_QR_NB ? k - j0 : NXC_QR_NB);
      s = nxc_qr_apply_block(ctx, qw, nq, j0, &vall, &tall, m, k, p, j0, nb, false);
    }
  }
  if (!s && n > 0) {
    nxc_qr_triu_kernel<T><<<dim3((unsigned)((m * n + 255) / 256), (unsigned)nbatch), 256, 0, ctx->stream>>>((T *)w->data, m, n);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "qr");
  }
  if (vbuf) nxc_free(ctx, vbuf);
  if (tbuf) nxc_free(ctx, tbuf);
  return s;
}

static bool nxc_qr_use_blocked(int cdt, int64_t m, int64_t n) {
  if (cdt != NXC_F32 && cdt != NXC_F64) return false;
  const int64_t k = m < n ? m : n;
  if (getenv("NX_CUDA_QR_BLOCKED")) return atoi(getenv("NX_CUDA_QR_BLOCKED")) != 0 && k > NXC_QR_NB;
  return k >= 128;
}

extern "C" nxc_status nxc_qr(nxc_ctx *ctx, const nxc_tensor *q, const nxc_tensor *r, const nxc_tensor *in, int reduced) {
  NXC_TRACE(ctx, "nxc_qr");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_qr (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(q)) || (s = nxc_check_tensor(r))) return nxc_la_fail(ctx, s);
  if (in->ndim < 2 || q->ndim != in->ndim || r->ndim != in->ndim) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t m = in->shape[in->ndim - 2], n = in->shape[in->ndim - 1];
  const int64_t k = m < n ? m : n;
  const int64_t nq = reduced ? k : m;
  if (q->shape[q->ndim - 2] != m || q->shape[q->ndim - 1] != nq) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  if (r->shape[r->ndim - 2] != nq || r->shape[r->ndim - 1] != n) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int cdt = nxc_la_compute_dtype(in->dtype);
  if (cdt < 0) return nxc_la_fail(ctx, NXC_LA_NOT_FLOAT);
  int64_t nbatch = 1;
  for (int i = 0; i < in->ndim - 2; i++) {
    if (q->shape[i] != in->shape[i] || r->shape[i] != in->shape[i]) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= in->shape[i];
  }
  if (nbatch == 0 || m == 0) return NXC_OK;
  nxc_tensor w, qw;
  void *wbuf = NULL, *qbuf = NULL, *tbuf = NULL;
  if ((s = nxc_la_work(ctx, in, cdt, m, n, &w, &wbuf))) return nxc_la_fail(ctx, s);
  s = nxc_la_work(ctx, in, cdt, m, nq, &qw, &qbuf);
  if (!s) s = nxc_alloc(ctx, (size_t)nbatch * (size_t)(k > 0 ? k : 1) * (size_t)nxc_elem_size(cdt), &tbuf);
  if (!s && n > 0) s = nxc_la_move(ctx, &w, in);
  if (!s && nxc_qr_use_blocked(cdt, m, n)) {
    s = cdt == NXC_F32 ? nxc_qr_blocked<float>(ctx, &w, &qw, m, n, nq, nbatch) : nxc_qr_blocked<double>(ctx, &w, &qw, m, n, nq, nbatch);
  } else if (!s) {
    NXC_LA_DISPATCH(cdt, { nxc_qr_kernel<T><<<(unsigned)nbatch, NXC_LA_THREADS, 0, ctx->stream>>>((T *)wbuf, (T *)qbuf, (T *)tbuf, m, n, nq); })
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "qr");
  }
  if (!s && nq > 0) s = nxc_la_move(ctx, q, &qw);
  if (!s && nq > 0 && n > 0) {
    // R = the first nq rows of the work matrix (its lower part already zeroed)
    nxc_tensor rv = w;
    rv.shape[rv.ndim - 2] = nq;
    s = nxc_la_move(ctx, r, &rv);
  }
  nxc_free(ctx, wbuf);
  if (qbuf) nxc_free(ctx, qbuf);
  if (tbuf) nxc_free(ctx, tbuf);
  return nxc_la_fail(ctx, s);
}

static const char NXC_LA_NO_CONVERGE[] = "eigenvalue iteration did not converge";

// eigh / eigvalsh (reference: caml_nx_c_eigh, nx_c_eigh.c; veneer backend_c/nx_backend.ml:627-648).
// `w` is f64 [batch, n]; `v` (input dtype, input shape) is written only when vectors != 0.
extern "C" nxc_status nxc_eigh(nxc_ctx *ctx, const nxc_tensor *w, const nxc_tensor *v, const nxc_tensor *in, int vectors) {
  NXC_TRACE(ctx, "nxc_eigh");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_eigh (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(w))) return nxc_la_fail(ctx, s);
  if (in->ndim < 2 || w->ndim != in->ndim - 1) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t n = in->shape[in->ndim - 1];
  if (in->shape[in->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_NOT_SQUARE);
  if (w->dtype != NXC_F64 || w->shape[w->ndim - 1] != n) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int cdt = nxc_la_compute_dtype(in->dtype);
  if (cdt < 0) return nxc_la_fail(ctx, NXC_LA_NOT_FLOAT);
  int64_t nbatch = 1;
  for (int i = 0; i < in->ndim - 2; i++) {
    if (w->shape[i] != in->shape[i]) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= in->shape[i];
  }
  if (vectors) {
    if ((s = nxc_check_tensor(v))) return nxc_la_fail(ctx, s);
    if (v->ndim != in->ndim || v->dtype != in->dtype) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    for (int i = 0; i < in->ndim; i++)
      if (v->shape[i] != in->shape[i]) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  }
  if (n == 0 || nbatch == 0) return NXC_OK;
  const size_t esz = (size_t)nxc_elem_size(cdt), nb = (size_t)nbatch, nn = (size_t)(n * n);
  // one allocation, carved: the input copy, G = A V, V, the sorted eigenvectors (compute type), w / sg (f64), rk, status
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_a = carve(nb * nn * esz), o_gt = carve(nb * nn * esz), o_vt = carve(nb * nn * esz);
  const size_t o_vo = carve(vectors ? nb * nn * esz : 16), o_w = carve(nb * n * 8), o_sg = carve(nb * n * 8);
  const size_t o_rk = carve(nb * n * 4), o_st = carve(sizeof(int)), o_team = carve(nb * NXC_LA_TEAM_BYTES);
  char *base = NULL;
  if ((s = nxc_alloc(ctx, off, (void **)&base))) return nxc_la_fail(ctx, s);
  s = nxc_memset(ctx, base + o_st, 0, sizeof(int));
  nxc_tensor aw = *in;
  aw.dtype = cdt; aw.offset = 0; aw.data = base + o_a;
  { int64_t stv = 1; for (int d = aw.ndim - 1; d >= 0; d--) { aw.strides[d] = stv; stv *= aw.shape[d]; } }
  if (!s) s = nxc_la_move(ctx, &aw, in);
  if (!s) {
    NxcEighArgs a;
    a.a = base + o_a; a.gt = base + o_gt; a.vt = base + o_vt; a.vo = base + o_vo;
    a.w = (double *)(base + o_w); a.sg = (double *)(base + o_sg); a.rk = (int *)(base + o_rk);
    a.n = n; a.vectors = vectors; a.status = (int *)(base + o_st); a.team = base + o_team;
    cudaError_t le = cudaSuccess;
    NXC_LA_DISPATCH(cdt, {
      auto kernel = nxc_eigh_kernel<typename La3Map<T>::E>;
      le = nxc_la_team_launch(ctx, kernel, nxc_la_team_size(kernel, nbatch, (n + 1) / 2), nbatch, a);
    })
    ctx->launches++;
    if (le != cudaSuccess) s = nxc_cuda_fail(ctx, le, "eigh");
    else if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "eigh");
  }
  if (!s) {
    int h = 0;
    NXC_CUDA_TRY(ctx, cudaMemcpyAsync(&h, base + o_st, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h == 3) s = NXC_LA_NO_CONVERGE;
  }
  if (!s) {
    nxc_tensor ww = *w;
    ww.offset = 0; ww.data = base + o_w;
    { int64_t stv = 1; for (int d = ww.ndim - 1; d >= 0; d--) { ww.strides[d] = stv; stv *= ww.shape[d]; } }
    s = nxc_copy(ctx, w, &ww);
  }
  if (!s && vectors) {
    nxc_tensor vw = aw;
    vw.data = base + o_vo;  // sorted eigenvectors
    s = nxc_la_move(ctx, v, &vw);
  }
  nxc_free(ctx, base);
  return nxc_la_fail(ctx, s);
}

// ---- linalg tier 3: svd, eig / eigvals ----------------------------------------------------------
// Kernel bodies: nxc_linalg3.cuh (shared with the CPU emulation the test suite runs).

struct NxcSvdArgs {
  void *gt, *wt, *ut, *uo, *vho;
  double *sig, *sg, *rown;
  int *rk;
  Cd *coef;
  int64_t m, n, ucols, vrows;
  int *status;
  char *team;  // NXC_LA_TEAM_BYTES per matrix when launched as clusters
};

template <class E>
__global__ void __launch_bounds__(NXC_LA_THREADS) nxc_svd_kernel(const NxcSvdArgs a) {
  __shared__ double sred[NXC_LA_THREADS / 32];
  __shared__ int sflags[2];
  unsigned rank, size;
  const La3Thr t = nxc_la3_team(&rank, &size);
  const int64_t b = blockIdx.x / size;
  double *red = t.wide ? (double *)(a.team + b * NXC_LA_TEAM_BYTES) : sred;
  int *flags = t.wide ? (int *)(a.team + b * NXC_LA_TEAM_BYTES + NXC_LA_TEAM_RED * 8) : sflags;
  const bool tall = a.m >= a.n;
  const int64_t pr = tall ? a.m : a.n, pc = tall ? a.n : a.m, ncu = tall ? a.ucols : a.vrows;
  la3_svd_body<E>(t, (E *)a.gt + b * pc * pr, (E *)a.wt + b * pc * pc, (E *)a.ut + b * ncu * pr,
                  (E *)a.uo + b * a.m * a.ucols, (E *)a.vho + b * a.vrows * a.n, a.sig + b * pc, a.sg + b * pc, a.rk + b * pc,
                  a.rown + b * pr, a.coef + b * pr, red, flags, a.m, a.n, a.ucols, a.vrows, 60, a.status);
}

struct NxcEigArgs {
  Cd *h, *z, *x, *vo, *w, *vs, *rs;
  double *rc, *bal;
  int *flag;
  int64_t n;
  int vectors;
  int smem_rows;
  int *status;
};

__global__ void __launch_bounds__(NXC_LA_THREADS, 2) nxc_eig_kernel(const __grid_constant__ NxcEigArgs a) {
  __shared__ double red[NXC_LA_THREADS];
  extern __shared__ __align__(16) unsigned char eig_smem[];
  const int64_t b = blockIdx.x, n = a.n, nn = a.n * a.n;
  // the Householder vector and the rotation chain are read by every thread at every step: keep them in
  // shared memory when they fit (44 bytes per row, with the wavefront's flags), in global scratch otherwise
  Cd *vs = a.smem_rows ? (Cd *)eig_smem : a.vs + b * n;
  Cd *rs = a.smem_rows ? vs + n : a.rs + b * n;
  double *rc = a.smem_rows ? (double *)(rs + n) : a.rc + b * n;
  int *flag = a.smem_rows ? (int *)(rc + n) : a.flag + b * n;
  la3_eig_body(nxc_la3_thr(), a.h + b * nn, a.z + b * nn, a.x + b * nn, a.vo + b * nn, a.w + b * n, vs, rc, rs, flag,
               a.bal + b * n, red, n, a.vectors, a.status);
}

// contiguous descriptor [batch..., rows, cols] of dtype dt over `data`, batch dims taken from `like`
static void nxc_la_desc(const nxc_tensor *like, int dt, int64_t rows, int64_t cols, void *data, nxc_tensor *w) {
  *w = *like;
  w->dtype = dt; w->offset = 0; w->data = data;
  w->shape[w->ndim - 2] = rows;
  w->shape[w->ndim - 1] = cols;
  int64_t st = 1;
  for (int d = w->ndim - 1; d >= 0; d--) { w->strides[d] = st; st *= w->shape[d]; }
}

static nxc_status nxc_la_status3(nxc_ctx *ctx, int *dev_status) {
  int h = 0;
  NXC_CUDA_TRY(ctx, cudaMemcpyAsync(&h, dev_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return h == 3 ? NXC_LA_NO_CONVERGE : NXC_OK;
}

// svd (reference: caml_nx_c_svd, nx_c_svd.c:2943-2958; driver nx_c_svd_run :2743-2938; veneer
// backend_c/nx_backend.ml:650-677). s is f64 [batch, k]; u [batch, m, k|m]; vt [batch, k|n, n]: thin or
// full is read off each output's shape, as the reference does.
extern "C" nxc_status nxc_svd(nxc_ctx *ctx, const nxc_tensor *u, const nxc_tensor *sv, const nxc_tensor *vt,
                              const nxc_tensor *in) {
  NXC_TRACE(ctx, "nxc_svd");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_svd (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(u)) || (s = nxc_check_tensor(sv)) || (s = nxc_check_tensor(vt)))
    return nxc_la_fail(ctx, s);
  if (in->ndim < 2 || u->ndim != in->ndim || vt->ndim != in->ndim || sv->ndim != in->ndim - 1)
    return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t m = in->shape[in->ndim - 2], n = in->shape[in->ndim - 1], k = m < n ? m : n;
  const int64_t ucols = u->shape[u->ndim - 1], vrows = vt->shape[vt->ndim - 2];
  if (u->shape[u->ndim - 2] != m || vt->shape[vt->ndim - 1] != n || sv->shape[sv->ndim - 1] != k ||
      (ucols != k && ucols != m) || (vrows != k && vrows != n))
    return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int cdt = nxc_la_compute_dtype(in->dtype);
  if (cdt < 0) return nxc_la_fail(ctx, NXC_LA_NOT_FLOAT);
  if (u->dtype != in->dtype || vt->dtype != in->dtype || sv->dtype != NXC_F64) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  int64_t nbatch = 1;
  for (int i = 0; i < in->ndim - 2; i++) {
    if (u->shape[i] != in->shape[i] || vt->shape[i] != in->shape[i] || sv->shape[i] != in->shape[i])
      return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= in->shape[i];
  }
  if (k == 0 || nbatch == 0) return NXC_OK;
  const bool tall = m >= n;
  const int64_t pr = tall ? m : n, pc = k, ncu = tall ? ucols : vrows;
  const size_t esz = (size_t)nxc_elem_size(cdt), nb = (size_t)nbatch;
  // one allocation, carved: gt, wt, ut, uo, vho (compute type), coef (Cd), sig, sg, rown (f64), rk (int), status
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_gt = carve(nb * pc * pr * esz), o_wt = carve(nb * pc * pc * esz), o_ut = carve(nb * ncu * pr * esz);
  const size_t o_uo = carve(nb * m * ucols * esz), o_vho = carve(nb * vrows * n * esz), o_coef = carve(nb * pr * sizeof(Cd));
  const size_t o_sig = carve(nb * pc * 8), o_sg = carve(nb * pc * 8), o_rown = carve(nb * pr * 8), o_rk = carve(nb * pc * 4);
  const size_t o_st = carve(sizeof(int)), o_team = carve(nb * NXC_LA_TEAM_BYTES);
  char *base = NULL;
  if ((s = nxc_alloc(ctx, off, (void **)&base))) return nxc_la_fail(ctx, s);
  s = nxc_memset(ctx, base + o_st, 0, sizeof(int));
  // gt = P^T without conjugation: the transposed view of A when tall, A itself when wide
  nxc_tensor gd, src = *in;
  nxc_la_desc(in, cdt, pc, pr, base + o_gt, &gd);
  if (tall) {
    const int r = in->ndim - 2, c = in->ndim - 1;
    src.shape[r] = in->shape[c]; src.shape[c] = in->shape[r];
    src.strides[r] = in->strides[c]; src.strides[c] = in->strides[r];
  }
  if (!s) s = nxc_la_move(ctx, &gd, &src);
  if (!s) {
    NxcSvdArgs a;
    a.gt = base + o_gt; a.wt = base + o_wt; a.ut = base + o_ut; a.uo = base + o_uo; a.vho = base + o_vho;
    a.sig = (double *)(base + o_sig); a.sg = (double *)(base + o_sg); a.rown = (double *)(base + o_rown);
    a.rk = (int *)(base + o_rk); a.coef = (Cd *)(base + o_coef);
    a.m = m; a.n = n; a.ucols = ucols; a.vrows = vrows; a.status = (int *)(base + o_st); a.team = base + o_team;
    cudaError_t le = cudaSuccess;
    NXC_LA_DISPATCH(cdt, {
      auto kernel = nxc_svd_kernel<typename La3Map<T>::E>;
      le = nxc_la_team_launch(ctx, kernel, nxc_la_team_size(kernel, nbatch, (k + 1) / 2), nbatch, a);
    })
    ctx->launches++;
    if (le != cudaSuccess) s = nxc_cuda_fail(ctx, le, "svd");
    else if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "svd");
  }
  if (!s) s = nxc_la_status3(ctx, (int *)(base + o_st));
  if (!s) {
    nxc_tensor d;
    nxc_la_desc(in, cdt, m, ucols, base + o_uo, &d);
    s = nxc_la_move(ctx, u, &d);
    if (!s) { nxc_la_desc(in, cdt, vrows, n, base + o_vho, &d); s = nxc_la_move(ctx, vt, &d); }
    if (!s) {
      nxc_tensor sd = *sv;
      sd.offset = 0; sd.data = base + o_sig;
      int64_t stv = 1;
      for (int dd = sd.ndim - 1; dd >= 0; dd--) { sd.strides[dd] = stv; stv *= sd.shape[dd]; }
      s = nxc_copy(ctx, sv, &sd);
    }
  }
  nxc_free(ctx, base);
  return nxc_la_fail(ctx, s);
}

static const char NXC_EIG_NOT_FLOAT[] = "eig requires a float or complex dtype";
#define NXC_EIG_MAX_N 46340  /* the reference's own bound (nx_c_eig.c:91-96) */
static const char NXC_EIG_TOO_LARGE[] = "matrix dimension exceeds eig limit";

// eig / eigvals (reference: caml_nx_c_eig, nx_c_eig.c:1310-1326; driver nx_c_eig_run :1195-1298; veneer
// backend_c/nx_backend.ml:679-707). w is c64 [batch, n]; v c64 [batch, n, n], written only when vectors != 0.
extern "C" nxc_status nxc_eig(nxc_ctx *ctx, const nxc_tensor *w, const nxc_tensor *v, const nxc_tensor *in, int vectors) {
  NXC_TRACE(ctx, "nxc_eig");
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_eig (reads a status word back)");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(w))) return nxc_la_fail(ctx, s);
  if (vectors && (s = nxc_check_tensor(v))) return nxc_la_fail(ctx, s);
  if (in->ndim < 2) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  const int64_t n = in->shape[in->ndim - 1];
  if (in->shape[in->ndim - 2] != n) return nxc_la_fail(ctx, NXC_LA_NOT_SQUARE);
  if (n > NXC_EIG_MAX_N) return nxc_la_fail(ctx, NXC_EIG_TOO_LARGE);
  if (w->ndim != in->ndim - 1 || w->shape[w->ndim - 1] != n) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  if (vectors && (v->ndim != in->ndim || v->shape[v->ndim - 1] != n || v->shape[v->ndim - 2] != n))
    return nxc_la_fail(ctx, NXC_LA_SHAPE);
  if (nxc_la_compute_dtype(in->dtype) < 0) return nxc_la_fail(ctx, NXC_EIG_NOT_FLOAT);
  if (w->dtype != NXC_C64 || (vectors && v->dtype != NXC_C64)) return nxc_la_fail(ctx, NXC_LA_SHAPE);
  int64_t nbatch = 1;
  for (int i = 0; i < in->ndim - 2; i++) {
    if (w->shape[i] != in->shape[i] || (vectors && v->shape[i] != in->shape[i])) return nxc_la_fail(ctx, NXC_LA_SHAPE);
    nbatch *= in->shape[i];
  }
  if (n == 0 || nbatch == 0) return NXC_OK;
  const size_t nb = (size_t)nbatch, nn = (size_t)(n * n);
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_h = carve(nb * nn * 16), o_z = carve(nb * nn * 16);
  const size_t o_x = carve(vectors ? nb * nn * 16 : 16), o_vo = carve(vectors ? nb * nn * 16 : 16);
  const size_t o_w = carve(nb * n * 16), o_vs = carve(nb * n * 16), o_rs = carve(nb * n * 16), o_rc = carve(nb * n * 8);
  const size_t o_bal = carve(nb * n * 8), o_flag = carve(nb * n * 4);
  const size_t o_st = carve(sizeof(int));
  char *base = NULL;
  if ((s = nxc_alloc(ctx, off, (void **)&base))) return nxc_la_fail(ctx, s);
  s = nxc_memset(ctx, base + o_st, 0, sizeof(int));
  nxc_tensor hd;
  nxc_la_desc(in, NXC_C64, n, n, base + o_h, &hd);
  if (!s) s = nxc_la_move(ctx, &hd, in);  // any float / complex storage -> complex double
  if (!s) {
    NxcEigArgs a;
    a.h = (Cd *)(base + o_h); a.z = (Cd *)(base + o_z); a.x = (Cd *)(base + o_x); a.vo = (Cd *)(base + o_vo);
    a.w = (Cd *)(base + o_w); a.vs = (Cd *)(base + o_vs); a.rs = (Cd *)(base + o_rs); a.rc = (double *)(base + o_rc); a.bal = (double *)(base + o_bal);
    a.flag = (int *)(base + o_flag);
    a.n = n; a.vectors = vectors; a.status = (int *)(base + o_st);
    size_t smem = (size_t)n * 44;
    a.smem_rows = smem <= 160 * 1024;
    if (!a.smem_rows) smem = 0;
    if (smem > 40 * 1024) cudaFuncSetAttribute(nxc_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nxc_eig_kernel<<<(unsigned)nbatch, NXC_LA_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "eig");
  }
  if (!s) s = nxc_la_status3(ctx, (int *)(base + o_st));
  if (!s) {
    nxc_tensor wd = *w;
    wd.offset = 0; wd.data = base + o_w;
    int64_t stv = 1;
    for (int dd = wd.ndim - 1; dd >= 0; dd--) { wd.strides[dd] = stv; stv *= wd.shape[dd]; }
    s = nxc_copy(ctx, w, &wd);
  }
  if (!s && vectors) {
    nxc_tensor vd;
    nxc_la_desc(in, NXC_C64, n, n, base + o_vo, &vd);
    s = nxc_copy(ctx, v, &vd);
  }
  nxc_free(ctx, base);
  return nxc_la_fail(ctx, s);
}
