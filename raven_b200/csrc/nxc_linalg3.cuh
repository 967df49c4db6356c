// nxc_linalg3.cuh -- linalg tier 3 kernel bodies: svd and eig / eigvals.
// Replaces caml_nx_c_svd (reference: nx_c_svd.c) and caml_nx_c_eig (reference: nx_c_eig.c).
//
// The reference's CONTRACTS are kept, not its algorithms (Golub-Kahan + divide-and-conquer for
// svd, EISPACK balanc/orthes/hqr2 for eig -- chains of scalar recurrences written for one core):
//   svd: A = U diag(S) V^H; S always float64, descending, non-negative; U / V^H orthonormal in the
//        input's compute type; thin (m x k, k x n) or full (m x m, n x n) by the output shapes
//        (backend_c/nx_backend.ml:650-677). Non-convergence is "eigenvalue iteration did not
//        converge" (nx_c_linalg.h LA_ERR_NO_CONVERGE), the string the veneer lifts to Linalg_error.
//   eig: eigenvalues and eigenvectors ALWAYS complex128 whatever the input dtype, eigenvector
//        columns of unit 2-norm, no phase or order convention (nx_c_eig.c:12-20, 59-65).
// How, GPU-shaped, one CTA per batch matrix:
//   svd: one-sided (Hestenes) Jacobi on the COLUMNS of the tall working matrix P (A, or A^H when
//        m < n), stored transposed so a column is a contiguous row. A round-robin schedule gives
//        pc/2 disjoint column pairs per step; a WARP owns a pair: three dot products (alpha, beta,
//        gamma, accumulated in double for every compute type, as the reference computes its
//        rotation scalars in double, nx_c_svd.c:37-39), then the rotation of the two columns and of
//        the matching columns of V -- no block barrier inside a step. Singular values are the final
//        column norms (high relative accuracy); zero columns (rank-deficient input) and the extra
//        columns of a full U / V are completed by two-pass Gram-Schmidt against the unit vector the
//        current basis covers least.
//   eig: everything in complex double (the output type): Householder reduction to Hessenberg form
//        with the transforms accumulated, then the explicitly shifted QR iteration (Wilkinson
//        shift, exceptional shifts at 10 / 20, a 30 n iteration cap -> no-convergence status, the
//        same discipline as nx_c_eig.c:44-52): the left Givens pass is a wavefront without
//        barriers (a thread carries its column down the chain and meets G_k as soon as the thread
//        that owns column k has published it through a flag), the right pass and the accumulation
//        into Z (kept transposed, so these row walks coalesce) are barrier-free too (a thread owns
//        a row and applies the whole rotation chain). Eigenvectors: one thread per eigenvalue back-substitutes on the
//        triangular Schur factor, then V = Z X and a column normalisation.
// The bodies are __host__ __device__ over an explicit thread descriptor so that tests/emu can run
// the very same code single-threaded on the CPU (tests only; the product has no CPU path).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef NXC_HD
#define NXC_HD __host__ __device__ __forceinline__
#endif

struct La3Thr {
  int tid, nt;       // thread index / count in the team working on one matrix
  int lane, lanes;   // lane in the warp / warp width (32 on the GPU, 1 in the emulation)
  int warp, nwarps;
  int wide;          // 0: the team is one CTA; 1: a thread-block cluster (red / flags then live in global memory)
};

// team barrier; for a cluster the release / acquire pair also orders the team's global-memory writes
NXC_HD void la3_sync(const La3Thr &t) {
#ifdef __CUDA_ARCH__
  if (t.wide) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
#else
  (void)t;
#endif
}
// the rotation counter is bumped by atomics from every CTA of the team: read it past the L1
NXC_HD int la3_flag(const int *flags) { return *(const volatile int *)flags; }
// one-way hand-off between threads of a CTA without a barrier: the producer's earlier stores are
// visible to a consumer that has seen the stamp (the emulation runs the producers first by construction)
NXC_HD void la3_publish(int *flag, int stamp) {
#ifdef __CUDA_ARCH__
  __threadfence_block();
  *(volatile int *)flag = stamp;
#else
  *flag = stamp;
#endif
}
// `far` != 0: this waiter is not next in line -- it backs off so that its polling does not take issue slots
// from the thread the chain is waiting for
NXC_HD void la3_await(const int *flag, int stamp, int far) {
#ifdef __CUDA_ARCH__
  while (*(const volatile int *)flag != stamp) {
    if (far) __nanosleep(200);
  }
  __threadfence_block();
#else
  (void)flag; (void)stamp; (void)far;
#endif
}
NXC_HD void la3_syncwarp() {
#ifdef __CUDA_ARCH__
  __syncwarp();
#endif
}
NXC_HD double la3_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
NXC_HD double la3_warp_sum(double v) {
#ifdef __CUDA_ARCH__
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
#endif
  return v;
}

struct Cd { double re, im; };
NXC_HD Cd cmk(double re, double im) { Cd z; z.re = re; z.im = im; return z; }
NXC_HD Cd cadd(Cd a, Cd b) { return cmk(a.re + b.re, a.im + b.im); }
NXC_HD Cd csub(Cd a, Cd b) { return cmk(a.re - b.re, a.im - b.im); }
NXC_HD Cd cmul(Cd a, Cd b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
NXC_HD Cd cscale(Cd a, double s) { return cmk(a.re * s, a.im * s); }
NXC_HD Cd cconj(Cd a) { return cmk(a.re, -a.im); }
NXC_HD double cnorm2(Cd a) { return a.re * a.re + a.im * a.im; }
NXC_HD double cabs1(Cd a) { return fabs(a.re) + fabs(a.im); }
NXC_HD double cabs_(Cd a) { return hypot(a.re, a.im); }
// Smith's division
NXC_HD Cd cdiv(Cd a, Cd b) {
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, d = b.re + b.im * r;
    return cmk((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  }
  const double r = b.re / b.im, d = b.re * r + b.im;
  return cmk((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
NXC_HD Cd csqrt_(Cd a) {
  const double m = cabs_(a);
  if (m == 0.0) return cmk(0.0, 0.0);
  const double t = sqrt(0.5 * (m + fabs(a.re)));
  if (a.re >= 0.0) return cmk(t, a.im / (2.0 * t));
  return cmk(fabs(a.im) / (2.0 * t), a.im >= 0.0 ? t : -t);
}

// element access of the four compute types as complex double
template <class T> struct La3El;
template <> struct La3El<float> {
  static const bool is_complex = false;
  NXC_HD static Cd ld(const float *p, int64_t i) { return cmk((double)p[i], 0.0); }
  NXC_HD static void st(float *p, int64_t i, Cd v) { p[i] = (float)v.re; }
  NXC_HD static double eps() { return 5.9604644775390625e-8; }
};
template <> struct La3El<double> {
  static const bool is_complex = false;
  NXC_HD static Cd ld(const double *p, int64_t i) { return cmk(p[i], 0.0); }
  NXC_HD static void st(double *p, int64_t i, Cd v) { p[i] = v.re; }
  NXC_HD static double eps() { return 1.1102230246251565e-16; }
};
struct La3C32 { float re, im; };
struct La3C64 { double re, im; };
template <> struct La3El<La3C32> {
  static const bool is_complex = true;
  NXC_HD static Cd ld(const La3C32 *p, int64_t i) { return cmk((double)p[i].re, (double)p[i].im); }
  NXC_HD static void st(La3C32 *p, int64_t i, Cd v) { p[i].re = (float)v.re; p[i].im = (float)v.im; }
  NXC_HD static double eps() { return 5.9604644775390625e-8; }
};
template <> struct La3El<La3C64> {
  static const bool is_complex = true;
  NXC_HD static Cd ld(const La3C64 *p, int64_t i) { return cmk(p[i].re, p[i].im); }
  NXC_HD static void st(La3C64 *p, int64_t i, Cd v) { p[i].re = v.re; p[i].im = v.im; }
  NXC_HD static double eps() { return 1.1102230246251565e-16; }
};

NXC_HD double la3_warp_max(double v) {
#ifdef __CUDA_ARCH__
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
#endif
  return v;
}
// every thread gets the CTA-wide sum (max) of v, combined in the same order everywhere: a shuffle
// butterfly inside each warp, then the per-warp partials through `red` (>= nwarps doubles)
NXC_HD double la3_block_sum(const La3Thr &t, double v, double *red) {
  v = la3_warp_sum(v);
  la3_sync(t);
  if (t.lane == 0) red[t.warp] = v;
  la3_sync(t);
  double s = 0.0;
  for (int i = 0; i < t.nwarps; i++) s += red[i];
  return s;
}
NXC_HD double la3_block_max(const La3Thr &t, double v, double *red) {
  v = la3_warp_max(v);
  la3_sync(t);
  if (t.lane == 0) red[t.warp] = v;
  la3_sync(t);
  double s = red[0];
  for (int i = 1; i < t.nwarps; i++) s = fmax(s, red[i]);
  return s;
}

// ---- svd ---------------------------------------------------------------------------------------
// Per batch matrix (all contiguous, compute type T unless stated):
//   gt  [pc][pr]  in: P^T without conjugation (row j = column j of A when m >= n, = row j of A
//                 when m < n: conjugated here since P = A^H); rotated in place
//   wt  [pc][pc]  V of P by columns-as-rows
//   ut  [ncu][pr] left singular vectors of P as rows, sorted, completed to ncu rows
//   uo  [m][ucols], vho [vrows][n]  the outputs in the compute type; ucols in {k, m}, vrows in {k, n}
//                 independently, as the reference reads them off the output shapes (nx_c_svd.c:2752-2757)
//   sig [pc] f64 sorted singular values; sg [pc] f64 / rk [pc] int / rown [pr] f64 / coef [pr] Cd scratch;
//   red: nt doubles (shared on the GPU); flags: 2 ints (shared): rotation count, spare
// status: 3 = did not converge.
template <class T>
NXC_HD void la3_svd_body(const La3Thr &t, T *gt, T *wt, T *ut, T *uo, T *vho, double *sig, double *sg, int *rk,
                         double *rown, Cd *coef, double *red, int *flags, int64_t m, int64_t n, int64_t ucols,
                         int64_t vrows, int max_sweeps, int *status) {
  typedef La3El<T> E;
  const bool tall = m >= n;
  const int64_t pr = tall ? m : n, pc = tall ? n : m;
  const int64_t ncu = tall ? ucols : vrows;  // left vectors of P wanted: pc (thin) or pr (full)
  // init: conjugate for the wide case, V = I
  if (!tall && E::is_complex)
    for (int64_t e = t.tid; e < pc * pr; e += t.nt) E::st(gt, e, cconj(E::ld(gt, e)));
  for (int64_t e = t.tid; e < pc * pc; e += t.nt) E::st(wt, e, cmk((e / pc) == (e % pc) ? 1.0 : 0.0, 0.0));
  la3_sync(t);
  const int64_t np_ = pc + (pc & 1), half = np_ / 2;
  const double tol = E::eps() * sqrt((double)pr);
  bool converged = pc <= 1;
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    if (t.tid == 0) flags[0] = 0;
    la3_sync(t);
    for (int64_t step = 0; step < np_ - 1; step++) {
      for (int64_t i = t.warp; i < half; i += t.nwarps) {
        // round-robin: player 0 fixed, the others rotate; pair i = (seat i, seat np-1-i)
        const int64_t k0 = i, k1 = np_ - 1 - i;
        int64_t p = k0 == 0 ? 0 : 1 + (k0 - 1 + (np_ - 1) - step) % (np_ - 1);
        int64_t q = k1 == 0 ? 0 : 1 + (k1 - 1 + (np_ - 1) - step) % (np_ - 1);
        if (p > q) { const int64_t s_ = p; p = q; q = s_; }
        if (q >= pc) continue;  // the padding player sits out
        T *x = gt + p * pr, *y = gt + q * pr;
        double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
        for (int64_t r = t.lane; r < pr; r += t.lanes) {
          const Cd a = E::ld(x, r), b = E::ld(y, r);
          al += cnorm2(a);
          be += cnorm2(b);
          gr += a.re * b.re + a.im * b.im;  // conj(a) * b
          gi += a.re * b.im - a.im * b.re;
        }
        al = la3_warp_sum(al); be = la3_warp_sum(be); gr = la3_warp_sum(gr); gi = la3_warp_sum(gi);
        const double g2 = gr * gr + gi * gi;
        if (!(al > 0.0) || !(be > 0.0) || !(g2 > tol * tol * al * be)) continue;
        const double ab = sqrt(g2);
        const double tau = (be - al) / (2.0 * ab);
        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = tt * c;
        const Cd su = cmk(s * gr / ab, s * gi / ab), suc = cconj(su);
        // x' = c x - s conj(u) y ; y' = s u x + c y   (columns of P and of V alike)
        for (int64_t r = t.lane; r < pr; r += t.lanes) {
          const Cd a = E::ld(x, r), b = E::ld(y, r);
          E::st(x, r, csub(cscale(a, c), cmul(suc, b)));
          E::st(y, r, cadd(cmul(su, a), cscale(b, c)));
        }
        T *vx = wt + p * pc, *vy = wt + q * pc;
        for (int64_t r = t.lane; r < pc; r += t.lanes) {
          const Cd a = E::ld(vx, r), b = E::ld(vy, r);
          E::st(vx, r, csub(cscale(a, c), cmul(suc, b)));
          E::st(vy, r, cadd(cmul(su, a), cscale(b, c)));
        }
        if (t.lane == 0) {
#ifdef __CUDA_ARCH__
          atomicAdd(&flags[0], 1);
#else
          flags[0] += 1;
#endif
        }
      }
      la3_sync(t);
    }
    if (la3_flag(flags) == 0) converged = true;
    la3_sync(t);
  }
  if (!converged) {
    // a last look with a looser bound before giving up: rounding can keep a pair hovering at tol
    if (t.tid == 0) flags[0] = 0;
    la3_sync(t);
    for (int64_t e = t.warp; e < pc * pc; e += t.nwarps) {
      const int64_t p = e / pc, q = e - p * pc;
      if (p >= q) continue;
      const T *x = gt + p * pr, *y = gt + q * pr;
      double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
      for (int64_t r = t.lane; r < pr; r += t.lanes) {
        const Cd a = E::ld(x, r), b = E::ld(y, r);
        al += cnorm2(a); be += cnorm2(b);
        gr += a.re * b.re + a.im * b.im;
        gi += a.re * b.im - a.im * b.re;
      }
      al = la3_warp_sum(al); be = la3_warp_sum(be); gr = la3_warp_sum(gr); gi = la3_warp_sum(gi);
      if (t.lane == 0 && al > 0.0 && be > 0.0 && gr * gr + gi * gi > 1e4 * tol * tol * al * be) flags[0] = 1;
    }
    la3_sync(t);
    if (la3_flag(flags) != 0) {
      if (t.tid == 0) {
#ifdef __CUDA_ARCH__
        atomicExch(status, 3);
#else
        *status = 3;
#endif
      }
      return;
    }
  }
  // singular values = column norms
  for (int64_t j = t.warp; j < pc; j += t.nwarps) {
    const T *x = gt + j * pr;
    double al = 0.0;
    for (int64_t r = t.lane; r < pr; r += t.lanes) al += cnorm2(E::ld(x, r));
    al = la3_warp_sum(al);
    if (t.lane == 0) sg[j] = sqrt(al);
  }
  la3_sync(t);
  // descending rank (ties by position); a column is "zero" when its norm vanishes against the largest
  // Jacobi keeps even rounding-level columns orthogonal to RELATIVE accuracy, so only a column that
  // is exactly zero (or whose squared norm underflows) has no direction of its own
  const double tiny = 0.0;
  for (int64_t j = t.tid; j < pc; j += t.nt) {
    const double sj = sg[j];
    int64_t rank = 0;
    for (int64_t k = 0; k < pc; k++) rank += (sg[k] > sj || (sg[k] == sj && k < j)) ? 1 : 0;
    rk[j] = (int)rank;
    sig[rank] = sj;
  }
  la3_sync(t);
  int64_t have = 0;  // columns with a usable direction: they sort first
  for (int64_t j = 0; j < pc; j++) have += (sg[j] > tiny) ? 1 : 0;
  // ut rows (sorted, normalised); rows >= have start as zero
  for (int64_t e = t.tid; e < ncu * pr; e += t.nt) E::st(ut, e, cmk(0.0, 0.0));
  la3_sync(t);
  for (int64_t j = t.warp; j < pc; j += t.nwarps) {
    if (!(sg[j] > tiny)) continue;
    const double inv = 1.0 / sg[j];
    const T *x = gt + j * pr;
    T *o = ut + (int64_t)rk[j] * pr;
    for (int64_t r = t.lane; r < pr; r += t.lanes) E::st(o, r, cscale(E::ld(x, r), inv));
  }
  la3_sync(t);
  if (have < ncu) {
    // complete: repeatedly orthogonalise the unit vector the basis covers least
    for (int64_t i = t.tid; i < pr; i += t.nt) {
      double s = 0.0;
      for (int64_t j = 0; j < have; j++) s += cnorm2(E::ld(ut, j * pr + i));
      rown[i] = s;
    }
    la3_sync(t);
    for (int64_t c = have; c < ncu; c++) {
      // argmin of rown (every thread scans: pr is small against the rest of the work)
      int64_t is = 0;
      double best = rown[0];
      for (int64_t i = 1; i < pr; i++)
        if (rown[i] < best) { best = rown[i]; is = i; }
      T *v = ut + c * pr;
      for (int64_t r = t.tid; r < pr; r += t.nt) {
        Cd acc = cmk(r == is ? 1.0 : 0.0, 0.0);
        for (int64_t j = 0; j < c; j++) acc = csub(acc, cmul(E::ld(ut, j * pr + r), cconj(E::ld(ut, j * pr + is))));
        E::st(v, r, acc);
      }
      la3_sync(t);
      for (int64_t j = t.warp; j < c; j += t.nwarps) {
        double cr = 0.0, ci = 0.0;
        for (int64_t r = t.lane; r < pr; r += t.lanes) {
          const Cd a = E::ld(ut, j * pr + r), b = E::ld(v, r);
          cr += a.re * b.re + a.im * b.im;
          ci += a.re * b.im - a.im * b.re;
        }
        cr = la3_warp_sum(cr); ci = la3_warp_sum(ci);
        if (t.lane == 0) coef[j] = cmk(cr, ci);
      }
      la3_sync(t);
      double nrm = 0.0;
      for (int64_t r = t.tid; r < pr; r += t.nt) {
        Cd acc = E::ld(v, r);
        for (int64_t j = 0; j < c; j++) acc = csub(acc, cmul(E::ld(ut, j * pr + r), coef[j]));
        E::st(v, r, acc);
        nrm += cnorm2(acc);
      }
      nrm = la3_block_sum(t, nrm, red);
      const double inv = 1.0 / sqrt(nrm);
      for (int64_t r = t.tid; r < pr; r += t.nt) {
        const Cd a = cscale(E::ld(v, r), inv);
        E::st(v, r, a);
        rown[r] += cnorm2(E::ld(v, r));
      }
      la3_sync(t);
    }
  }
  // outputs
  if (tall) {
    for (int64_t e = t.tid; e < m * ucols; e += t.nt) {
      const int64_t r = e / ucols, j = e - r * ucols;
      E::st(uo, e, E::ld(ut, j * pr + r));
    }
    for (int64_t e = t.tid; e < pc * pc; e += t.nt) {
      const int64_t j = e / pc, c = e - j * pc;  // source row j of wt = column j of V
      E::st(vho, (int64_t)rk[j] * n + c, cconj(E::ld(wt, e)));
    }
  } else {
    for (int64_t e = t.tid; e < pc * pc; e += t.nt) {
      const int64_t j = e / pc, r = e - j * pc;
      E::st(uo, r * ucols + rk[j], E::ld(wt, e));
    }
    for (int64_t e = t.tid; e < vrows * n; e += t.nt) E::st(vho, e, cconj(E::ld(ut, e)));
  }
}

// ---- eigh ----------------------------------------------------------------------------------------
// Hermitian eigendecomposition (reference: nx_c_eigh.c tridiagonalises and runs implicit-shift QL on
// one core; its contract: LOWER triangle read, eigenvalues ascending and always f64, orthonormal
// eigenvector columns in the input's compute type). Jacobi again, in its ONE-SIDED (implicit) form so
// that it shares the svd kernel's shape: keep G = A V and V, both by columns-as-rows; the pivot block
// a two-sided sweep would read off V^H A V is three dot products, a_pp = v_p . g_p, a_qq = v_q . g_q,
// a_pq = v_p . g_q, and the rotation touches only columns p and q of G and V -- contiguous rows
// here, a WARP per pair, no block barrier inside a round-robin step. (The two-sided kernel this
// replaces rotated columns of a row-major matrix, 32 sectors per warp access: 4.5 s for one 512 x 512
// matrix against 0.5 s for the svd of the same size.) lambda_j = v_j . g_j at the end.
//   a  [n][n] in: the matrix as moved in (row-major; only its lower triangle is read)
//   gt [n][n], vt [n][n] work; vo [n][n] out: eigenvectors (columns), sorted; w [n] f64 out
//   sg [n] f64 / rk [n] int scratch; red: nt doubles; flags: 2 ints.   status 3 = did not converge.
template <class T>
NXC_HD void la3_eigh_body(const La3Thr &t, const T *a, T *gt, T *vt, T *vo, double *w, double *sg, int *rk, double *red,
                          int *flags, int64_t n, int vectors, int max_sweeps, int *status) {
  typedef La3El<T> E;
  double part = 0.0;
  for (int64_t e = t.tid; e < n * n; e += t.nt) {
    const int64_t j = e / n, r = e - j * n;        // gt[j][r] = A[r][j]
    Cd v = j <= r ? E::ld(a, r * n + j) : cconj(E::ld(a, j * n + r));
    if (j == r) v.im = 0.0;
    E::st(gt, e, v);
    E::st(vt, e, cmk(j == r ? 1.0 : 0.0, 0.0));
    part += cnorm2(v);
  }
  const double norm2 = la3_block_sum(t, part, red);
  la3_sync(t);
  const int64_t np_ = n + (n & 1), half = np_ / 2;
  // rotate while |a_pq| stands out of the rounding of its own dot product: eps * ||A||_F
  const double thr2 = E::eps() * E::eps() * norm2;
  bool converged = n <= 1 || !(norm2 > 0.0);
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    if (t.tid == 0) flags[0] = 0;
    la3_sync(t);
    for (int64_t step = 0; step < np_ - 1; step++) {
      for (int64_t i = t.warp; i < half; i += t.nwarps) {
        const int64_t k0 = i, k1 = np_ - 1 - i;
        int64_t p = k0 == 0 ? 0 : 1 + (k0 - 1 + (np_ - 1) - step) % (np_ - 1);
        int64_t q = k1 == 0 ? 0 : 1 + (k1 - 1 + (np_ - 1) - step) % (np_ - 1);
        if (p > q) { const int64_t s_ = p; p = q; q = s_; }
        if (q >= n) continue;
        T *gp = gt + p * n, *gq = gt + q * n, *vp = vt + p * n, *vq = vt + q * n;
        double app = 0.0, aqq = 0.0, gr = 0.0, gi = 0.0;
        for (int64_t r = t.lane; r < n; r += t.lanes) {
          const Cd xp = E::ld(vp, r), xq = E::ld(vq, r), yp = E::ld(gp, r), yq = E::ld(gq, r);
          app += xp.re * yp.re + xp.im * yp.im;
          aqq += xq.re * yq.re + xq.im * yq.im;
          gr += xp.re * yq.re + xp.im * yq.im;  // conj(v_p) * g_q
          gi += xp.re * yq.im - xp.im * yq.re;
        }
        app = la3_warp_sum(app); aqq = la3_warp_sum(aqq); gr = la3_warp_sum(gr); gi = la3_warp_sum(gi);
        const double g2 = gr * gr + gi * gi;
        if (!(g2 > thr2)) continue;
        const double ab = sqrt(g2);
        const double tau = (aqq - app) / (2.0 * ab);
        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = tt * c;
        const Cd su = cmk(s * gr / ab, s * gi / ab), suc = cconj(su);
        for (int64_t r = t.lane; r < n; r += t.lanes) {
          const Cd xp = E::ld(gp, r), xq = E::ld(gq, r);
          E::st(gp, r, csub(cscale(xp, c), cmul(suc, xq)));
          E::st(gq, r, cadd(cmul(su, xp), cscale(xq, c)));
          const Cd zp = E::ld(vp, r), zq = E::ld(vq, r);
          E::st(vp, r, csub(cscale(zp, c), cmul(suc, zq)));
          E::st(vq, r, cadd(cmul(su, zp), cscale(zq, c)));
        }
        if (t.lane == 0) {
#ifdef __CUDA_ARCH__
          atomicAdd(&flags[0], 1);
#else
          flags[0] += 1;
#endif
        }
      }
      la3_sync(t);
    }
    if (la3_flag(flags) == 0) converged = true;
    la3_sync(t);
  }
  if (!converged) {
    // accept what is diagonal to 100 eps ||A||_F: rounding can keep a pair hovering at the threshold
    if (t.tid == 0) flags[0] = 0;
    la3_sync(t);
    for (int64_t e = t.warp; e < n * n; e += t.nwarps) {
      const int64_t p = e / n, q = e - p * n;
      if (p >= q) continue;
      const T *vp = vt + p * n, *gq = gt + q * n;
      double gr = 0.0, gi = 0.0;
      for (int64_t r = t.lane; r < n; r += t.lanes) {
        const Cd xp = E::ld(vp, r), yq = E::ld(gq, r);
        gr += xp.re * yq.re + xp.im * yq.im;
        gi += xp.re * yq.im - xp.im * yq.re;
      }
      gr = la3_warp_sum(gr); gi = la3_warp_sum(gi);
      if (t.lane == 0 && gr * gr + gi * gi > 1e4 * thr2) flags[0] = 1;
    }
    la3_sync(t);
    if (la3_flag(flags) != 0) {
      if (t.tid == 0) {
#ifdef __CUDA_ARCH__
        atomicExch(status, 3);
#else
        *status = 3;
#endif
      }
      return;
    }
  }
  // Rayleigh quotients, ascending rank (ties by position), sorted outputs
  for (int64_t j = t.warp; j < n; j += t.nwarps) {
    const T *vj = vt + j * n, *gj = gt + j * n;
    double lam = 0.0;
    for (int64_t r = t.lane; r < n; r += t.lanes) {
      const Cd x = E::ld(vj, r), y = E::ld(gj, r);
      lam += x.re * y.re + x.im * y.im;
    }
    lam = la3_warp_sum(lam);
    if (t.lane == 0) sg[j] = lam;
  }
  la3_sync(t);
  for (int64_t j = t.tid; j < n; j += t.nt) {
    const double lj = sg[j];
    int64_t rank = 0;
    for (int64_t k = 0; k < n; k++) rank += (sg[k] < lj || (sg[k] == lj && k < j)) ? 1 : 0;
    rk[j] = (int)rank;
    w[rank] = lj;
  }
  la3_sync(t);
  if (vectors)
    for (int64_t e = t.tid; e < n * n; e += t.nt) {
      const int64_t j = e / n, r = e - j * n;
      E::st(vo, r * n + rk[j], E::ld(vt, e));
    }
}

// ---- eig -----------------------------------------------------------------------------------------
// Per batch matrix, all complex double, contiguous:
//   h [n][n] in: the matrix; out: the triangular Schur factor      z [n][n] accumulated unitary
//   x [n][n] scratch (triangular eigenvectors)                       vo [n][n] out: eigenvectors (columns)
//   w [n] out: eigenvalues     vs [n] Cd, rc [n] double, rs [n] Cd, bal [n] double scratch;  red: nt doubles
// status: 3 = did not converge.
NXC_HD void la3_eig_body(const La3Thr &t, Cd *h, Cd *z, Cd *x, Cd *vo, Cd *w, Cd *vs, double *rc, Cd *rs, int *flag,
                         double *bal, double *red, int64_t n, int vectors, int *status) {
  const double eps = 2.220446049250313e-16;
  // Z is kept TRANSPOSED (z[k * n + i] = Z[i][k]): every pass over Z gives a thread a ROW of Z and walks its
  // columns, so consecutive threads then touch consecutive addresses
  for (int64_t e = t.tid; e < n * n; e += t.nt) z[e] = cmk((e / n) == (e % n) ? 1.0 : 0.0, 0.0);
  for (int64_t i = t.tid; i < n; i += t.nt) { bal[i] = 1.0; flag[i] = 0; }
  la3_sync(t);
  // Balancing, the scaling half of the reference's `balanc` (nx_c_eig.c:25-27; EISPACK balanc / LAPACK
  // gebal without the permutation phase): a diagonal similarity D^-1 A D by exact powers of two that
  // brings each row's and column's 1-norm together, so the QR iteration's eps * ||A|| errors are not
  // dominated by a few huge entries. Sequential in i by nature; the two norms are CTA-wide sums.
  for (int pass = 0; pass < 16; pass++) {
    bool changed = false;
    for (int64_t i = 0; i < n; i++) {
      double cp = 0.0, rp = 0.0;
      for (int64_t j = t.tid; j < n; j += t.nt)
        if (j != i) { cp += cabs1(h[j * n + i]); rp += cabs1(h[i * n + j]); }
      double c = la3_block_sum(t, cp, red);
      const double r = la3_block_sum(t, rp, red);
      if (!(c > 0.0) || !(r > 0.0) || !(c < 1e300) || !(r < 1e300)) continue;  // uniform: every thread holds the same sums
      double f = 1.0;
      const double s = c + r;
      double g = r * 0.5;
      while (c < g) { f *= 2.0; c *= 4.0; }
      g = r * 2.0;
      while (c >= g) { f *= 0.5; c *= 0.25; }
      if ((c + r) / f < 0.95 * s) {
        changed = true;
        la3_sync(t);  // everyone has read the old sums' inputs
        const double gi = 1.0 / f;
        if (t.tid == 0) bal[i] *= f;
        for (int64_t j = t.tid; j < n; j += t.nt) {
          if (j == i) continue;
          h[i * n + j] = cscale(h[i * n + j], gi);
          h[j * n + i] = cscale(h[j * n + i], f);
        }
        la3_sync(t);  // ... and the next row's sums see the scaled entries
      }
    }
    la3_sync(t);
    if (!changed) break;
  }
  la3_sync(t);
  // Householder reduction to upper Hessenberg form: H <- Q^H H Q, Z <- Z Q  (zlarfg convention:
  // Q = I - tau v v^H with v[0] = 1, Q^H x = beta e1, beta real)
  for (int64_t k = 0; k + 2 < n; k++) {
    double part = 0.0;
    for (int64_t i = k + 2 + t.tid; i < n; i += t.nt) part += cnorm2(h[i * n + k]);
    const double xnorm2 = la3_block_sum(t, part, red);
    const Cd alpha = h[(k + 1) * n + k];
    if (xnorm2 == 0.0 && alpha.im == 0.0) continue;  // uniform across the CTA
    const double nrm = sqrt(cnorm2(alpha) + xnorm2);
    const double beta = alpha.re >= 0.0 ? -nrm : nrm;
    const Cd tau = cmk((beta - alpha.re) / beta, -alpha.im / beta);
    const Cd scal = cdiv(cmk(1.0, 0.0), cmk(alpha.re - beta, alpha.im));
    la3_sync(t);
    for (int64_t i = k + 1 + t.tid; i < n; i += t.nt) {
      vs[i] = i == k + 1 ? cmk(1.0, 0.0) : cmul(h[i * n + k], scal);
      h[i * n + k] = i == k + 1 ? cmk(beta, 0.0) : cmk(0.0, 0.0);
    }
    la3_sync(t);
    const Cd tauc = cconj(tau);
    // left: H <- (I - conj(tau) v v^H) H on columns k+1..n-1
    for (int64_t j = k + 1 + t.tid; j < n; j += t.nt) {
      Cd d = cmk(0.0, 0.0);
      for (int64_t i = k + 1; i < n; i++) d = cadd(d, cmul(cconj(vs[i]), h[i * n + j]));
      d = cmul(d, tauc);
      for (int64_t i = k + 1; i < n; i++) h[i * n + j] = csub(h[i * n + j], cmul(vs[i], d));
    }
    la3_sync(t);
    // right: H <- H (I - tau v v^H), Z likewise; a thread owns a row
    // (rows of H are contiguous: a WARP takes a row, lanes along it; rows of Z are columns of the transposed
    // storage: a THREAD takes a row, consecutive threads side by side -- both coalesced)
    // (small matrices: a row is shorter than the warp's reduction, a thread per row of H is quicker)
    const bool wide_rows = n >= 96;
    for (int64_t i = t.tid; !wide_rows && i < n; i += t.nt) {
      Cd *row = h + i * n;
      Cd d = cmk(0.0, 0.0);
      for (int64_t j = k + 1; j < n; j++) d = cadd(d, cmul(row[j], vs[j]));
      d = cmul(d, tau);
      for (int64_t j = k + 1; j < n; j++) row[j] = csub(row[j], cmul(d, cconj(vs[j])));
    }
    for (int64_t i = t.warp; wide_rows && i < n; i += t.nwarps) {
      Cd *row = h + i * n;
      double dr = 0.0, di = 0.0;
      for (int64_t j = k + 1 + t.lane; j < n; j += t.lanes) {
        const Cd p = cmul(row[j], vs[j]);
        dr += p.re; di += p.im;
      }
      const Cd d = cmul(cmk(la3_warp_sum(dr), la3_warp_sum(di)), tau);
      for (int64_t j = k + 1 + t.lane; j < n; j += t.lanes) row[j] = csub(row[j], cmul(d, cconj(vs[j])));
    }
    for (int64_t i = t.tid; i < n; i += t.nt) {
      Cd *row = z + i;
      Cd d = cmk(0.0, 0.0);
      for (int64_t j = k + 1; j < n; j++) d = cadd(d, cmul(row[j * n], vs[j]));
      d = cmul(d, tau);
      for (int64_t j = k + 1; j < n; j++) row[j * n] = csub(row[j * n], cmul(d, cconj(vs[j])));
    }
    la3_sync(t);
  }
  // a scale for the deflation test when both neighbours vanish
  double part = 0.0;
  for (int64_t e = t.tid; e < n * n; e += t.nt) part += cabs1(h[e]);
  const double hnorm = la3_block_sum(t, part, red);
  // shifted QR on the active block [l, hi]
  int64_t hi = n - 1;
  int iter = 0;
  int64_t total = 0;
  const int64_t cap = 30 * n + 30;
  while (hi >= 0) {
    // the active block starts at the LARGEST l <= hi whose subdiagonal entry is negligible (0 if none):
    // every thread tests a strided share of the candidates, one CTA-wide max -- a serial walk down the
    // subdiagonal is a chain of dependent L2 round trips on every QR step
    double cand = 0.0;
    for (int64_t i = hi - t.tid; i > 0; i -= t.nt) {
      double tst = cabs1(h[(i - 1) * n + (i - 1)]) + cabs1(h[i * n + i]);
      if (tst == 0.0) tst = hnorm;
      if (cabs1(h[i * n + (i - 1)]) <= eps * tst) { cand = (double)i; break; }  // this thread's largest
    }
    const int64_t l = (int64_t)la3_block_max(t, cand, red);
    la3_sync(t);  // everyone has read the subdiagonal before it is cleaned
    if (l > 0 && t.tid == 0) h[l * n + (l - 1)] = cmk(0.0, 0.0);
    la3_sync(t);
    if (l == hi) {
      if (t.tid == 0) w[hi] = h[hi * n + hi];
      hi--;
      iter = 0;
      la3_sync(t);
      continue;
    }
    if (total >= cap) {
      if (t.tid == 0) {
#ifdef __CUDA_ARCH__
        atomicExch(status, 3);
#else
        *status = 3;
#endif
      }
      return;
    }
    // Wilkinson shift: the eigenvalue of the trailing 2x2 nearer to its last diagonal entry
    Cd mu;
    {
      const Cd a = h[(hi - 1) * n + (hi - 1)], b = h[(hi - 1) * n + hi], c = h[hi * n + (hi - 1)], d = h[hi * n + hi];
      if (iter == 10 || iter == 20) {
        mu = cmk(d.re + 0.75 * cabs1(c) + (hi >= 2 ? 0.4375 * cabs1(h[(hi - 1) * n + (hi - 2)]) : 0.0), d.im);
      } else {
        const Cd tr2 = cscale(csub(a, d), 0.5);
        const Cd bc = cmul(b, c);
        Cd sq = csqrt_(cadd(cmul(tr2, tr2), bc));
        if (tr2.re * sq.re + tr2.im * sq.im < 0.0) sq = cmk(-sq.re, -sq.im);
        const Cd den = cadd(tr2, sq);
        mu = cnorm2(den) > 0.0 ? csub(d, cdiv(bc, den)) : d;
      }
    }
    la3_sync(t);
    for (int64_t k = l + t.tid; k <= hi; k += t.nt) h[k * n + k] = csub(h[k * n + k], mu);
    la3_sync(t);
    // left pass: R = G_{hi-1} ... G_l (H - mu I) as a WAVEFRONT. A thread owns the columns j = l + tid
    // (mod nt), in increasing order, and carries each down the chain: rows k, k+1 of column j meet
    // G_k = [c s; -conj(s) c] as soon as the owner of column k has published it (flag[k] = this sweep's
    // stamp) -- the owner being the thread that has just brought column k down to its diagonal. One
    // step of the chain costs a rotation's arithmetic and a flag, not a CTA barrier (a barrier per
    // column was 0.5 ms per sweep of a 512-wide block), and the threads to the right trail the
    // diagonal by one step each. Nobody but its owner touches a column during the pass.
    const int stamp = (int)(total + 1);
    for (int64_t j = l + t.tid; j < n; j += t.nt) {
      const int64_t kend = j < hi ? j : hi;
      Cd u0 = h[l * n + j];
      // the subdiagonal entry G_j is made from: asked for now, its L2 round trip is over when the chain arrives
      const Cd b0 = j < hi ? h[(j + 1) * n + j] : cmk(0.0, 0.0);
      for (int64_t k = l; k < kend; k++) {
        const Cd u1 = h[(k + 1) * n + j];
        la3_await(&flag[k], stamp, j - k > 48);
        const double c = rc[k];
        const Cd s = rs[k];
        h[k * n + j] = cadd(cscale(u0, c), cmul(s, u1));
        u0 = csub(cscale(u1, c), cmul(cconj(s), u0));
      }
      if (j < hi) {
        // (these few flops sit on the critical path of every column: one scaling division and two
        // square roots instead of three hypot calls)
        Cd a = u0, b = b0;
        double c = 1.0;
        Cd s = cmk(0.0, 0.0);
        const double sc = fmax(cabs1(a), cabs1(b));
        if (sc > 0.0 && !(b.re == 0.0 && b.im == 0.0)) {
          // the scaling division only when the squares could leave the double range; the two reciprocal
          // square roots are independent of each other (the chain waits for this thread)
          if (!(sc > 1e-140 && sc < 1e140)) {
            const double isc = 1.0 / sc;
            a = cscale(a, isc); b = cscale(b, isc);
          }
          const double na2 = cnorm2(a), nb2 = cnorm2(b);
          if (na2 == 0.0) { c = 0.0; s = cscale(cconj(b), 1.0 / sqrt(nb2)); }
          else {
            const double ina = la3_rsqrt(na2), ir = la3_rsqrt(na2 + nb2);
            c = na2 * ina * ir;
            s = cscale(cmul(a, cconj(b)), ir * ina);
          }
        }
        rc[j] = c; rs[j] = s;
        la3_publish(&flag[j], stamp);
        h[j * n + j] = cadd(cscale(u0, c), cmul(s, b0));
        h[(j + 1) * n + j] = cmk(0.0, 0.0);
      } else {
        h[kend * n + j] = u0;
      }
    }
    la3_sync(t);
    // right pass: H <- R G_l^H ... G_{hi-1}^H and Z likewise; a thread owns a row and walks the chain
    for (int64_t i = t.tid; i <= hi + n; i += t.nt) {
      const bool isz = i > hi;
      Cd *row = isz ? z + (i - hi - 1) : h + i * n;
      const int64_t st = isz ? n : 1;
      int64_t k0 = l;
      if (!isz && i - 1 > l) k0 = i - 1;
      if (isz && !vectors) continue;
      Cd u0 = row[k0 * st];
      for (int64_t k = k0; k < hi; k++) {
        const double c = rc[k];
        const Cd s = rs[k], sc = cconj(s);
        const Cd u1 = row[(k + 1) * st];
        row[k * st] = cadd(cscale(u0, c), cmul(u1, sc));
        u0 = csub(cscale(u1, c), cmul(u0, s));
      }
      if (k0 < hi) row[hi * st] = u0;
    }
    la3_sync(t);
    for (int64_t k = l + t.tid; k <= hi; k += t.nt) h[k * n + k] = cadd(h[k * n + k], mu);
    la3_sync(t);
    iter++;
    total++;
  }
  la3_sync(t);
  if (!vectors) return;
  // eigenvectors of the triangular factor: thread k back-substitutes column k
  const double smin = eps * (hnorm > 0.0 ? hnorm / (double)n : 1.0);
  for (int64_t k = t.tid; k < n; k += t.nt) {
    const Cd lam = h[k * n + k];
    for (int64_t i = n - 1; i > k; i--) x[i * n + k] = cmk(0.0, 0.0);
    x[k * n + k] = cmk(1.0, 0.0);
    for (int64_t i = k - 1; i >= 0; i--) {
      Cd s = cmk(0.0, 0.0);
      for (int64_t j = i + 1; j <= k; j++) s = cadd(s, cmul(h[i * n + j], x[j * n + k]));
      Cd d = csub(h[i * n + i], lam);
      if (cabs1(d) < smin) d = cmk(smin, 0.0);
      const Cd xi = cdiv(cmk(-s.re, -s.im), d);
      x[i * n + k] = xi;
      // a (nearly) defective matrix grows the column by 1/smin per step: rescale before it overflows
      // (an eigenvector's scale is free; the column is normalised below)
      const double big = cabs1(xi);
      if (big > 1e150) {
        const double inv = 1.0 / big;
        for (int64_t j = i; j <= k; j++) x[j * n + k] = cscale(x[j * n + k], inv);
      }
    }
  }
  la3_sync(t);
  // V = Z X (X upper triangular), then unit 2-norm columns
  for (int64_t e = t.tid; e < n * n; e += t.nt) {
    const int64_t r = e / n, k = e - r * n;
    Cd acc = cmk(0.0, 0.0);
    for (int64_t j = 0; j <= k; j++) acc = cadd(acc, cmul(z[j * n + r], x[j * n + k]));
    vo[e] = acc;
  }
  la3_sync(t);
  for (int64_t e = t.tid; e < n * n; e += t.nt) vo[e] = cscale(vo[e], bal[e / n]);  // undo the balancing: v = D y
  la3_sync(t);
  for (int64_t k = t.tid; k < n; k += t.nt) {
    double big = 0.0;
    for (int64_t r = 0; r < n; r++) { const double a = cabs1(vo[r * n + k]); big = a > big ? a : big; }
    if (!(big > 0.0)) continue;
    double s = 0.0;
    for (int64_t r = 0; r < n; r++) s += cnorm2(cscale(vo[r * n + k], 1.0 / big));
    const double inv = 1.0 / (big * sqrt(s));
    for (int64_t r = 0; r < n; r++) vo[r * n + k] = cscale(vo[r * n + k], inv);
  }
}
