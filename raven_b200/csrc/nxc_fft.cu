// nxc_fft.cu -- fft / ifft / rfft / irfft.
// Replaces caml_nx_c_fft / caml_nx_c_ifft / caml_nx_c_rfft / caml_nx_c_irfft (reference:
// nx_c_fft.c:1173-1223; drivers nx_c_fft.c:940-1143; per-axis pass nx_c_fft.c:869-936).
//
// What is kept from the reference is the CONTRACT: transforms are unnormalised (fft is the
// -sign DFT, ifft the +sign DFT without 1/n; nx_c_fft.c:38-41), everything computes in double
// (c32 / f32 upcast on gather and round once on store), multi-axis = one 1-D pass per axis in
// the reference's order (fft: axes order, first pass reads `in`, later ones `out` in place;
// rfft: the last axis first, real -> half spectrum, then the others; irfft: the other axes in
// a c64 temporary, then half spectrum -> real with Hermitian reconstruction and zero-padding /
// truncation to `s`), any length n is accepted.
//
// How it is computed is not the reference's (a CPU mixed-radix Cooley-Tukey with per-thread
// line scratch).  Powers of two up to 4096^2 points never touch a gather / scatter pass:
//   lines    nxc_fft_lines_kernel: a CTA reads whole lines straight from the source tensor
//            (strided, converted to double complex, Hermitian-mirrored for irfft), runs every
//            stage in shared memory (two radix-2 stages per pass in registers, stage-major
//            twiddle table, swizzled slots) and writes straight to the destination tensor
//   > 4096   four-step through the same kernel: columns of the m1 x m2 view into a work buffer,
//            then twiddled rows with the transposed write (nxc_fft_pow2)
//   other n  Bluestein's chirp-z (m = next power of two >= 2n-1): gather with the chirp multiply
//            into a contiguous [lines][m] work buffer, the two forward legs and the inverse leg on
//            the kernels above, pointwise product with the transformed filter, scatter with the
//            chirp multiply and 1/m
//   > 2^24   log2(m) global radix-2 passes (kept as the fallback)
// Twiddles are sincospi() of an exactly reduced rational angle in double precision, tabulated
// once per call.
#include "nxc_common.cuh"
#include "nxc_map.cuh"

#define NXC_FFT_SMEM_MAX 4096

struct NxcFftLines {
  int n;  // dims other than the transform axis
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t shape[NXC_MAX_NDIM];
  int64_t src_stride[NXC_MAX_NDIM];
  int64_t dst_stride[NXC_MAX_NDIM];
  int small;
};

__device__ __forceinline__ void nxc_fft_line_base(const NxcFftLines &d, int64_t L, int64_t &so, int64_t &dof) {
  so = 0;
  dof = 0;
  if (d.small) {
    uint32_t r = (uint32_t)L;
    for (int i = d.n - 1; i >= 0; i--) {
      const uint32_t q = nxc_fastdiv(r, d.div[i]);
      const uint32_t c = r - q * d.div[i].d;
      so += (int64_t)c * d.src_stride[i];
      dof += (int64_t)c * d.dst_stride[i];
      r = q;
    }
  } else {
    int64_t r = L;
    for (int i = d.n - 1; i >= 0; i--) {
      const int64_t q = r / d.shape[i];
      const int64_t c = r - q * d.shape[i];
      so += c * d.src_stride[i];
      dof += c * d.dst_stride[i];
      r = q;
    }
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// exp(sign * i * pi * num / den), 0 <= num < 2*den
__device__ __forceinline__ double2 cispi(int sign, int64_t num, int64_t den) {
  double s, c;
  sincospi((double)num / (double)den, &s, &c);
  return make_double2(c, sign < 0 ? -s : s);
}
// Bluestein chirp w[k] = exp(sign * i * pi * k^2 / n), k^2 reduced mod 2n (reference:
// nx_c_fft.c:656-663)
__device__ __forceinline__ double2 chirp_at(int sign, int64_t k, int64_t n) {
  const unsigned long long kk = ((unsigned long long)k * (unsigned long long)k) % (unsigned long long)(2 * n);
  return cispi(sign, (int64_t)kk, n);
}

enum { NXC_FFT_SRC_C32 = 0, NXC_FFT_SRC_C64 = 1, NXC_FFT_SRC_F32 = 2, NXC_FFT_SRC_F64 = 3 };

struct NxcFftGather {
  NxcFftLines lines;
  const void *src;
  double2 *work;
  int64_t n_lines;
  int64_t n;        // logical line length fed to the transform
  int64_t n_src;    // elements present in the source line (irfft: usable half-spectrum bins)
  int64_t m;        // work line length (n, or Bluestein's power of two)
  int64_t stride;   // source stride along the axis (elements)
  int src_kind;
  int sign;
  int bluestein;
  int hermitian;    // irfft last axis: rebuild the length-n spectrum from n_src bins
  const double2 *filter;  // Bluestein's second leg: element k is multiplied by filter[k] on the way in
  const double2 *chirp;   // tabulated chirp_at(sign, k, n), k < n (NULL: computed per element)
};

__device__ __forceinline__ double2 nxc_fft_read(const void *src, int kind, int64_t off) {
  switch (kind) {
    case NXC_FFT_SRC_C32: { const float2 z = ((const float2 *)src)[off]; return make_double2((double)z.x, (double)z.y); }
    case NXC_FFT_SRC_C64: return ((const double2 *)src)[off];
    case NXC_FFT_SRC_F32: return make_double2((double)((const float *)src)[off], 0.0);
    default: return make_double2(((const double *)src)[off], 0.0);
  }
}

// element k of line L's transform input (line base `so` in source elements)
__device__ __forceinline__ double2 nxc_fft_fetch(const NxcFftGather &g, int64_t so, int64_t k) {
  double2 v = make_double2(0.0, 0.0);
  if (k < g.n) {
    if (k < g.n_src) {
      v = nxc_fft_read(g.src, g.src_kind, so + k * g.stride);
    } else if (g.hermitian) {
      // conjugate mirror of bin n-k when that bin was supplied (reference: nx_c_fft.c:1014-1021)
      const int64_t q = g.n - k;
      if (q >= 1 && q < g.n_src && q != k) {
        v = nxc_fft_read(g.src, g.src_kind, so + q * g.stride);
        v.y = -v.y;
      }
    }
    if (g.bluestein) v = cmul(v, g.chirp ? __ldg(g.chirp + k) : chirp_at(g.sign, k, g.n));
    if (g.filter) v = cmul(v, __ldg(g.filter + k));
  }
  return v;
}

__global__ void __launch_bounds__(256) nxc_fft_gather_kernel(const __grid_constant__ NxcFftGather g) {
  const int64_t total = g.n_lines * g.m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t L = i / g.m, k = i - L * g.m;
    int64_t so = 0, dof = 0;
    if (k < g.n) nxc_fft_line_base(g.lines, L, so, dof);
    g.work[i] = nxc_fft_fetch(g, so, k);
  }
}

struct NxcFftScatter {
  NxcFftLines lines;
  void *dst;
  const double2 *work;
  int64_t n_lines;
  int64_t n, n_out, m;
  int64_t stride;  // destination stride along the axis
  int dst_kind;
  int sign;
  int bluestein;
  const double2 *chirp;  // as NxcFftGather::chirp
};

// output element k of a line (line base `dof` in destination elements)
__device__ __forceinline__ void nxc_fft_put(const NxcFftScatter &g, int64_t dof, int64_t k, double2 v) {
  if (g.bluestein) {
    v = cmul(v, g.chirp ? __ldg(g.chirp + k) : chirp_at(g.sign, k, g.n));
    const double inv = 1.0 / (double)g.m;
    v.x *= inv;
    v.y *= inv;
  }
  const int64_t off = dof + k * g.stride;
  switch (g.dst_kind) {
    case NXC_FFT_SRC_C32: ((float2 *)g.dst)[off] = make_float2((float)v.x, (float)v.y); break;
    case NXC_FFT_SRC_C64: ((double2 *)g.dst)[off] = v; break;
    case NXC_FFT_SRC_F32: ((float *)g.dst)[off] = (float)v.x; break;
    default: ((double *)g.dst)[off] = v.x; break;
  }
}

__global__ void __launch_bounds__(256) nxc_fft_scatter_kernel(const __grid_constant__ NxcFftScatter g) {
  const int64_t total = g.n_lines * g.n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t L = i / g.n_out, k = i - L * g.n_out;
    int64_t so, dof;
    nxc_fft_line_base(g.lines, L, so, dof);
    nxc_fft_put(g, dof, k, g.work[L * g.m + k]);
  }
}

// ---- the power-of-two core -------------------------------------------------------------------
// Stockham radix-2, stage Ns = 1, 2, 4, ...: butterfly j (0 <= j < m/2) reads in[j], in[j + m/2],
// twiddles the second by exp(sign*i*pi*(j mod Ns)/Ns), and writes a+b, a-b to
// out[(j div Ns)*2Ns + j mod Ns] and Ns further: the output of the last stage is in natural order.
//
// Lines of m <= NXC_FFT_SMEM_MAX points never leave shared memory: a CTA takes `lpc` whole lines,
// reads them STRAIGHT from the source tensor (FUSED: strided, converted, Hermitian-mirrored as the
// gather kernel would) or from the work buffer, runs every stage between two ping-pong buffers and
// writes the result straight to the destination -- the transform costs one read and one write of
// the data.  Two radix-2 stages are taken per pass in registers (nxc_fft_stages: a 1024-point
// line is 5 shared-memory round trips, not 10) between two ping-pong buffers, and twiddles come from a table computed once per call with sincospi of the
// reduced angle -- the value each butterfly used to recompute in double precision.
// tw[Ns + k] = exp(sign i pi k / Ns) for Ns = 1, 2, 4, ..., m / 2 and k < Ns: every stage reads a
// CONTIGUOUS run (one table indexed k * (m / 2 / Ns) made a warp touch 32 cache lines per load)
__global__ void __launch_bounds__(256) nxc_fft_twiddle_kernel(double2 *tw, int64_t m, int sign) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == 0) { tw[0] = make_double2(1.0, 0.0); continue; }
    const int64_t Ns = (int64_t)1 << (63 - __clzll(i));
    tw[i] = cispi(sign, i - Ns, Ns);
  }
}

// Lines longer than NXC_FFT_SMEM_MAX take the four-step route through the same kernel: a line of
// M = m1 m2 points is an m1 x m2 matrix; MODE_COLS transforms its m2 columns (length m1, read from
// the source tensor, written in place into a work buffer), MODE_ROWS multiplies element (k1, n2) by
// exp(sign 2 pi i k1 n2 / M), transforms the m1 rows (length m2) and writes row k1's bin k2 to output
// bin k1 + m1 k2 -- two round trips through memory where log2(M) radix-2 passes took one each.
enum { NXC_FFT_MODE_LINE = 0, NXC_FFT_MODE_COLS = 1, NXC_FFT_MODE_ROWS = 2 };
struct NxcFftCore {
  double2 *work;        // four-step intermediate: [outer lines][M], written by COLS, read by ROWS
  const double2 *tw;    // m twiddles, stage-major (nxc_fft_twiddle_kernel)
  const double2 *t_lo, *t_hi;  // ROWS: exp(sign 2 pi i e / M) = t_hi[e >> lo_bits] t_lo[e & mask]
  int sign;
  int64_t n_lines;      // lines this launch transforms (outer lines x `other` in the four-step modes)
  int m, log2m, lpc;
  int npass, plan;      // stages taken by pass i = (plan >> 4 i) & 15
  int swz;              // slot swizzle shift (>= 3)
  int mode, log2other, lo_bits;
  int lfast_in, lfast_out;  // MODE_LINE along a strided axis: walk the CTA's lines fastest (adjacent lines are adjacent in memory)
};

// Shared-memory slot of element o of a line: the first pass makes thread j write the 2^R contiguous
// elements 2^R j .., one per store instruction -- unswizzled, the threads of a quarter-warp would
// land on one or two 16-byte bank groups (4-way conflicts for R = 2).  XOR with the bits above an
// aligned group of 8 permutes slots inside that group, so the contiguous runs every other access
// makes stay conflict-free too.
__device__ __forceinline__ int nxc_fft_swz(int o, int sh) { return o ^ ((o >> sh) & 7); }

// R radix-2 stages on the 2^R points x[j + t m / 2^R] held in registers: after r stages the points
// form 2^r sets (set q continues at output offset q Ns with twiddle index k + q Ns), and stage r
// turns set q into sets q (sums) and q + 2^r (differences) -- the arithmetic of R Stockham passes,
// one shared-memory round trip.  Twiddles of the upper half of a stage's sets are the lower half's
// turned by a quarter (exact), so a pass loads 2^(R-1) of them, each a contiguous run across threads.
template <int R>
__device__ __forceinline__ void nxc_fft_stages(double2 (&v)[1 << R], const double2 *__restrict__ tw, int k, int Ns, int sign) {
  constexpr int N = 1 << R;
#pragma unroll
  for (int r = 0; r < R; r++) {
    double2 u[N], w[N / 2 > 0 ? N / 2 : 1];
#pragma unroll
    for (int q = 0; q < (1 << r); q++) {
      if (r == 0 || q < (1 << (r - (r > 0)))) {
        w[q] = __ldg(tw + (Ns << r) + k + q * Ns);
      } else {
        const double2 z = w[q - (1 << (r - (r > 0)))];
        w[q] = sign < 0 ? make_double2(z.y, -z.x) : make_double2(-z.y, z.x);
      }
#pragma unroll
      for (int t = 0; t < (N >> (r + 1)); t++) {
        const double2 a = v[q + (t << r)];
        const double2 b = cmul(v[q + (t << r) + (N >> 1)], w[q]);
        u[q + (t << (r + 1))] = make_double2(a.x + b.x, a.y + b.y);
        u[q + (1 << r) + (t << (r + 1))] = make_double2(a.x - b.x, a.y - b.y);
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = u[i];
  }
}

// one pass over the CTA's lines, buffer `src` to buffer `dst`, a group of 2^R points at a time
template <int R>
__device__ __forceinline__ void nxc_fft_pass_r(const double2 *src, double2 *dst, const NxcFftCore &c, int nl, int st) {
  constexpr int N = 1 << R;
  const int groups = c.m >> R, Ns = 1 << st;
  for (int i = threadIdx.x; i < nl * groups; i += blockDim.x) {
    const int l = i >> (c.log2m - R), j = i & (groups - 1), k = j & (Ns - 1);
    const double2 *x = src + (size_t)l * c.m;
    double2 *y = dst + (size_t)l * c.m;
    double2 v[N];
#pragma unroll
    for (int t = 0; t < N; t++) v[t] = x[nxc_fft_swz(j + t * groups, c.swz)];
    nxc_fft_stages<R>(v, c.tw, k, Ns, c.sign);
    const int o = ((j - k) << R) + k;
#pragma unroll
    for (int q = 0; q < N; q++) y[nxc_fft_swz(o + q * Ns, c.swz)] = v[q];
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(1024) nxc_fft_lines_kernel(const __grid_constant__ NxcFftCore c, const __grid_constant__ NxcFftGather g,
                                                             const __grid_constant__ NxcFftScatter sc) {
  extern __shared__ __align__(16) unsigned char nxc_fft_smem[];
  const int m = c.m;
  double2 *a = (double2 *)nxc_fft_smem, *b = a + (size_t)c.lpc * m;
  int64_t *base = (int64_t *)(b + (size_t)c.lpc * m);  // [lpc][2]: source / destination bases of the OUTER line
  const int64_t L0 = (int64_t)blockIdx.x * c.lpc;
  const int nl = (int)(c.n_lines - L0 < c.lpc ? c.n_lines - L0 : c.lpc);
  const int log2other = MODE == NXC_FFT_MODE_LINE ? 0 : c.log2other;
  const int other = 1 << log2other;                // sub-lines per outer line
  const int64_t M = (int64_t)m << log2other;
  for (int l = threadIdx.x; l < nl; l += blockDim.x)
    nxc_fft_line_base(g.lines, (L0 + l) >> log2other, base[2 * l], base[2 * l + 1]);
  __syncthreads();
  // load: along the line fastest, except for columns (adjacent lines are adjacent in memory)
  for (int i = threadIdx.x; i < nl * m; i += blockDim.x) {
    int l, k;
    if (MODE == NXC_FFT_MODE_COLS || (MODE == NXC_FFT_MODE_LINE && c.lfast_in)) { k = i / nl; l = i - k * nl; }
    else { l = i >> c.log2m; k = i & (m - 1); }
    const int64_t L = L0 + l;
    const int sub = (int)(L & (other - 1));
    double2 v;
    if (MODE == NXC_FFT_MODE_ROWS) {
      v = c.work[(L >> log2other) * M + (int64_t)sub * m + k];
      const int64_t e = (int64_t)sub * k;
      v = cmul(v, cmul(__ldg(c.t_hi + (e >> c.lo_bits)), __ldg(c.t_lo + (e & (((int64_t)1 << c.lo_bits) - 1)))));
    } else {
      v = nxc_fft_fetch(g, base[2 * l], MODE == NXC_FFT_MODE_COLS ? (int64_t)k * other + sub : (int64_t)k);
    }
    a[(size_t)l * m + nxc_fft_swz(k, c.swz)] = v;
  }
  __syncthreads();
  int st = 0;
  for (int pi = 0; pi < c.npass; pi++) {
    const int R = (c.plan >> (4 * pi)) & 15;
    switch (R) {
      case 2: nxc_fft_pass_r<2>(a, b, c, nl, st); break;
      default: nxc_fft_pass_r<1>(a, b, c, nl, st); break;
    }
    { double2 *t = a; a = b; b = t; }
    st += R;
  }
  if (MODE == NXC_FFT_MODE_LINE) {
    const int n_out = (int)(sc.n_out < m ? sc.n_out : m);
    for (int i = threadIdx.x; i < nl * n_out; i += blockDim.x) {
      int l, k;
      if (c.lfast_out) { k = i / nl; l = i - k * nl; }
      else { l = i / n_out; k = i - l * n_out; }
      nxc_fft_put(sc, base[2 * l + 1], k, a[(size_t)l * m + nxc_fft_swz(k, c.swz)]);
    }
  } else {
    for (int i = threadIdx.x; i < nl * m; i += blockDim.x) {
      const int k = i / nl, l = i - k * nl;  // adjacent lines are adjacent outputs in both modes
      const int64_t L = L0 + l;
      const int sub = (int)(L & (other - 1));
      const double2 v = a[(size_t)l * m + nxc_fft_swz(k, c.swz)];
      if (MODE == NXC_FFT_MODE_COLS) {
        c.work[(L >> log2other) * M + (int64_t)k * other + sub] = v;
      } else {
        const int64_t bin = sub + (int64_t)other * k;
        if (bin < sc.n_out) nxc_fft_put(sc, base[2 * l + 1], bin, v);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
nxc_fft_global_pass_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, int64_t n_lines, int64_t m,
                           int st, int sign) {
  const int64_t half = m >> 1, total = n_lines * half;
  const int64_t Ns = (int64_t)1 << st;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t L = i / half, j = i - L * half;
    const double2 *a = in + L * m;
    double2 *b = out + L * m;
    const int64_t k = j & (Ns - 1);
    const double2 u0 = a[j];
    const double2 u1 = cmul(a[j + half], cispi(sign, k, Ns));
    const int64_t j0 = ((j - k) << 1) + k;
    b[j0] = make_double2(u0.x + u1.x, u0.y + u1.y);
    b[j0 + Ns] = make_double2(u0.x - u1.x, u0.y - u1.y);
  }
}

// Bluestein filter b[k] = conj(w[k]) for |k| < n, circular over m, zero elsewhere
__global__ void __launch_bounds__(256) nxc_fft_filter_kernel(double2 *bf, int64_t n, int64_t m, int sign) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    int64_t q = -1;
    if (k < n) q = k;
    else if (m - k < n) q = m - k;
    if (q >= 0) { v = chirp_at(sign, q, n); v.y = -v.y; }
    bf[k] = v;
  }
}
__global__ void __launch_bounds__(256) nxc_fft_chirp_kernel(double2 *ch, int64_t n, int sign) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    ch[k] = chirp_at(sign, k, n);
}
__global__ void __launch_bounds__(256)
nxc_fft_pointwise_kernel(double2 *work, const double2 *__restrict__ bf, int64_t n_lines, int64_t m) {
  const int64_t total = n_lines * m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    work[i] = cmul(work[i], bf[i % m]);
}

static unsigned nxc_fft_grid(nxc_ctx *ctx, int64_t items) {
  int64_t b = (items + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static int nxc_fft_log2(int64_t m) {
  int l = 0;
  while (((int64_t)1 << l) < m) l++;
  return l;
}

// One launch of the lines kernel: `n_lines` lines of m <= NXC_FFT_SMEM_MAX points in `mode`.
static nxc_status nxc_fft_lines_launch(nxc_ctx *ctx, NxcFftCore c, const NxcFftGather &g, const NxcFftScatter &sc) {
  const int64_t m = c.m;
  c.log2m = nxc_fft_log2(m);
  double2 *tw = NULL;
  nxc_status s = nxc_alloc(ctx, sizeof(double2) * (size_t)m, (void **)&tw);
  if (s) return s;
  nxc_fft_twiddle_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(tw, m, c.sign);
  ctx->launches++;
  c.tw = tw;
  // stages per pass: nxc_fft_stages<R> is written for any R, and R = 3 / 4 were measured (16384 lines of 1024
  // points: 0.26 / 0.29 ms at 76 / 128 registers against 0.22 ms for R = 2 at 48): the transform is latency-
  // bound, so 40 resident warps per SM beat fewer shared-memory round trips
  const int rmax = 2;
  int rmin = rmax;
  {
    int rem = c.log2m, first = rem < rmax ? rem : rmax;
    c.npass = 0; c.plan = 0;
    c.plan |= first << (4 * c.npass++);
    c.swz = first < 3 ? 3 : first;
    rmin = first;
    rem -= first;
    const int cnt = (rem + rmax - 1) / rmax;
    for (int i = 0; i < cnt; i++) {
      const int r = (rem + cnt - 1 - i) / cnt;
      c.plan |= r << (4 * c.npass++);
      if (r < rmin) rmin = r;
    }
  }
  // lines per CTA: enough for one group of 2^R points per thread in the pass with the most groups; the
  // four-step modes want at least two (their strided side then moves whole 32-byte sectors)
  const int64_t gmax = m >> rmin;
  c.lpc = (int)(gmax >= 256 ? 1 : 256 / gmax);
  if (c.mode != NXC_FFT_MODE_LINE && c.lpc < 2 && m <= NXC_FFT_SMEM_MAX / 2) c.lpc = 2;
  if (c.mode == NXC_FFT_MODE_LINE) {
    c.lfast_in = g.stride != 1 && g.stride != -1;
    c.lfast_out = sc.stride != 1 && sc.stride != -1;
    // a strided axis: as many adjacent lines per CTA as the two buffers hold, up to a 128-byte run
    if (c.lfast_in || c.lfast_out)
      while (c.lpc < 8 && (int64_t)c.lpc * 2 * m <= NXC_FFT_SMEM_MAX) c.lpc *= 2;
  }
  if (c.lpc > c.n_lines) c.lpc = (int)c.n_lines;
  int64_t want = gmax * c.lpc;
  const int threads = want <= 256 ? 256 : (want >= 1024 ? 1024 : (int)want);
  const size_t smem = (size_t)c.lpc * (2 * (size_t)m * sizeof(double2) + 16);
  void (*kernel)(const NxcFftCore, const NxcFftGather, const NxcFftScatter) =
      c.mode == NXC_FFT_MODE_LINE ? nxc_fft_lines_kernel<NXC_FFT_MODE_LINE>
                                  : (c.mode == NXC_FFT_MODE_COLS ? nxc_fft_lines_kernel<NXC_FFT_MODE_COLS>
                                                                 : nxc_fft_lines_kernel<NXC_FFT_MODE_ROWS>);
  // per device, not per process: a flag remembered across contexts would skip the second device of a process
  static bool attr_set[3][64] = {};
  const int dev = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
  if (!attr_set[c.mode][dev] || ctx->device >= 64) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * NXC_FFT_SMEM_MAX * sizeof(double2) + 16 * 256));
    if (e != cudaSuccess) { nxc_free(ctx, tw); return nxc_cuda_fail(ctx, e, "fft attribute"); }
    attr_set[c.mode][dev] = true;
  }
  const int64_t ctas = (c.n_lines + c.lpc - 1) / c.lpc;
  if (ctas > 0x7FFFFFFF) { nxc_free(ctx, tw); return NXC_ERR_SHAPE; }  // > 2^31 CTAs: beyond any device's memory
  kernel<<<(unsigned)ctas, threads, smem, ctx->stream>>>(c, g, sc);
  ctx->launches++;
  if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "fft lines");
  nxc_free(ctx, tw);
  return s;
}

__global__ void __launch_bounds__(256) nxc_fft_split_twiddle_kernel(double2 *t_lo, double2 *t_hi, int lo_bits, int hi_bits,
                                                                    int64_t M, int sign) {
  const int64_t nlo = (int64_t)1 << lo_bits, nhi = (int64_t)1 << hi_bits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nlo + nhi; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nlo) t_lo[i] = cispi(sign, 2 * i, M);
    else t_hi[i - nlo] = cispi(sign, 2 * ((i - nlo) << lo_bits), M);
  }
}

// Power-of-two transform of `n_outer` lines of m points described by a gather / scatter pair (any
// source / destination the descriptors can address): one launch up to NXC_FFT_SMEM_MAX points, the
// four-step pair of launches through `work` ([n_outer][m] double2) up to NXC_FFT_SMEM_MAX^2.
static nxc_status nxc_fft_pow2(nxc_ctx *ctx, const NxcFftGather &g, const NxcFftScatter &sc, int64_t n_outer, int64_t m,
                               int sign, double2 *work) {
  NxcFftCore c;
  memset(&c, 0, sizeof c);
  c.sign = sign;
  if (m <= NXC_FFT_SMEM_MAX) {
    c.mode = NXC_FFT_MODE_LINE; c.m = (int)m; c.n_lines = n_outer; c.log2other = 0;
    return nxc_fft_lines_launch(ctx, c, g, sc);
  }
  const int log2M = nxc_fft_log2(m), l1 = log2M / 2, l2 = log2M - l1;  // columns of 2^l1 points, rows of 2^l2
  double2 *tt = NULL;
  const int lo_bits = (log2M + 1) / 2, hi_bits = log2M - lo_bits;
  nxc_status s = nxc_alloc(ctx, sizeof(double2) * (((size_t)1 << lo_bits) + ((size_t)1 << hi_bits)), (void **)&tt);
  if (s) return s;
  nxc_fft_split_twiddle_kernel<<<64, 256, 0, ctx->stream>>>(tt, tt + ((size_t)1 << lo_bits), lo_bits, hi_bits, m, sign);
  ctx->launches++;
  c.work = work;
  c.mode = NXC_FFT_MODE_COLS; c.m = 1 << l1; c.log2other = l2; c.n_lines = n_outer << l2;
  s = nxc_fft_lines_launch(ctx, c, g, sc);
  if (!s) {
    c.mode = NXC_FFT_MODE_ROWS; c.m = 1 << l2; c.log2other = l1; c.n_lines = n_outer << l1;
    c.t_lo = tt; c.t_hi = tt + ((size_t)1 << lo_bits); c.lo_bits = lo_bits;
    s = nxc_fft_lines_launch(ctx, c, g, sc);
  }
  nxc_free(ctx, tt);
  return s;
}

// gather / scatter descriptors of n_lines contiguous double-complex lines of m points at `buf`
static void nxc_fft_contig(double2 *buf, int64_t n_lines, int64_t m, NxcFftGather *g, NxcFftScatter *sc) {
  memset(g, 0, sizeof *g);
  memset(sc, 0, sizeof *sc);
  NxcFftLines ln;
  memset(&ln, 0, sizeof ln);
  ln.n = 1;
  ln.shape[0] = n_lines; ln.src_stride[0] = m; ln.dst_stride[0] = m;
  ln.small = n_lines < 0x7FFFFFFFLL;
  ln.div[0] = nxc_fastdiv_make(ln.small ? (uint32_t)n_lines : 1u);
  g->lines = ln; g->src = buf; g->n_lines = n_lines; g->n = m; g->n_src = m; g->m = m; g->stride = 1;
  g->src_kind = NXC_FFT_SRC_C64;
  sc->lines = ln; sc->dst = buf; sc->n_lines = n_lines; sc->n = m; sc->n_out = m; sc->m = m; sc->stride = 1;
  sc->dst_kind = NXC_FFT_SRC_C64;
}

// In-place transform of n_lines contiguous lines of power-of-two length m (the Bluestein legs).
// `tmp` is a second buffer of the same size, used when m > NXC_FFT_SMEM_MAX.
static nxc_status nxc_fft_core(nxc_ctx *ctx, double2 *work, double2 *tmp, int64_t n_lines, int64_t m, int sign) {
  const int log2m = nxc_fft_log2(m);
  if (log2m == 0) return NXC_OK;
  if (m <= (int64_t)NXC_FFT_SMEM_MAX * NXC_FFT_SMEM_MAX) {
    NxcFftGather g;
    NxcFftScatter sc;
    nxc_fft_contig(work, n_lines, m, &g, &sc);
    // long lines: columns go work -> tmp, rows tmp -> work (each CTA's reads and writes are the same elements
    // only in the single-launch case, where a line is wholly in shared memory before it is written)
    return nxc_fft_pow2(ctx, g, sc, n_lines, m, sign, tmp);
  }
  double2 *a = work, *b = tmp;
  for (int st = 0; st < log2m; st++) {
    nxc_fft_global_pass_kernel<<<nxc_fft_grid(ctx, n_lines * (m / 2)), 256, 0, ctx->stream>>>(a, b, n_lines, m, st, sign);
    NXC_LAUNCH_CHECK(ctx);
    double2 *t = a; a = b; b = t;
  }
  if (a != work) NXC_CUDA_TRY(ctx, cudaMemcpyAsync(work, a, sizeof(double2) * (size_t)(n_lines * m), cudaMemcpyDeviceToDevice, ctx->stream));
  return NXC_OK;
}

struct NxcFftPass {
  const nxc_tensor *src, *dst;
  int src_kind, dst_kind;
  int axis;
  int64_t n;       // transform length
  int64_t n_src;   // source bins read per line
  int64_t n_out;   // elements written per line
  int sign;
  int hermitian;
};

static int64_t nxc_fft_esize(int kind) { return kind == NXC_FFT_SRC_C64 ? 16 : (kind == NXC_FFT_SRC_F32 ? 4 : 8); }

// One axis pass over all lines (reference: run_axis, nx_c_fft.c:896-936).
static nxc_status nxc_fft_pass(nxc_ctx *ctx, const NxcFftPass &p) {
  const nxc_tensor *src = p.src, *dst = p.dst;
  int64_t n_lines = 1;
  NxcFftLines ln;
  ln.n = 0;
  for (int d = 0; d < dst->ndim; d++) {
    if (d == p.axis) continue;
    n_lines *= dst->shape[d];
    ln.shape[ln.n] = dst->shape[d];
    ln.src_stride[ln.n] = src->strides[d];
    ln.dst_stride[ln.n] = dst->strides[d];
    ln.n++;
  }
  if (n_lines == 0 || p.n == 0 || p.n_out == 0) return NXC_OK;
  ln.small = n_lines < 0x7FFFFFFFLL;
  for (int i = 0; i < ln.n; i++) ln.div[i] = nxc_fastdiv_make(ln.small ? (uint32_t)ln.shape[i] : 1u);
  const bool pow2 = (p.n & (p.n - 1)) == 0;
  int64_t m = p.n;
  if (!pow2) {
    m = 1;
    while (m < 2 * p.n - 1) m <<= 1;
  }
  if (pow2 && m >= 2 && m <= (int64_t)NXC_FFT_SMEM_MAX * NXC_FFT_SMEM_MAX) {
    // source tensor -> shared memory -> destination tensor, once (twice through a work buffer for long lines)
    NxcFftGather g;
    g.lines = ln;
    g.src = (const char *)src->data + src->offset * nxc_fft_esize(p.src_kind);
    g.work = NULL;
    g.n_lines = n_lines; g.n = p.n; g.n_src = p.n_src; g.m = m;
    g.stride = src->strides[p.axis];
    g.src_kind = p.src_kind; g.sign = p.sign; g.bluestein = 0; g.hermitian = p.hermitian; g.filter = NULL; g.chirp = NULL;
    NxcFftScatter sc;
    sc.lines = ln;
    sc.dst = (char *)dst->data + dst->offset * nxc_fft_esize(p.dst_kind);
    sc.work = NULL;
    sc.n_lines = n_lines; sc.n = p.n; sc.n_out = p.n_out; sc.m = m;
    sc.stride = dst->strides[p.axis];
    sc.dst_kind = p.dst_kind; sc.sign = p.sign; sc.bluestein = 0; sc.chirp = NULL;
    double2 *mid = NULL;
    nxc_status s4 = NXC_OK;
    if (m > NXC_FFT_SMEM_MAX) s4 = nxc_alloc(ctx, sizeof(double2) * (size_t)n_lines * (size_t)m, (void **)&mid);
    if (!s4) s4 = nxc_fft_pow2(ctx, g, sc, n_lines, m, p.sign, mid);
    if (mid) nxc_free(ctx, mid);
    return s4;
  }
  if (!pow2 && m <= (int64_t)NXC_FFT_SMEM_MAX * NXC_FFT_SMEM_MAX) {
    // Bluestein in two transforms: (source x chirp, zero-padded) -> work, then (work x filter) -> inverse
    // transform -> (x chirp / m) -> destination; the chirp, the filter product and the final scaling ride on
    // the lines kernel's loads and stores
    double2 *work = NULL, *mid = NULL, *bf = NULL;
    nxc_status sb = nxc_alloc(ctx, sizeof(double2) * (size_t)n_lines * (size_t)m, (void **)&work);
    if (!sb && m > NXC_FFT_SMEM_MAX) sb = nxc_alloc(ctx, sizeof(double2) * (size_t)n_lines * (size_t)m, (void **)&mid);
    if (!sb) sb = nxc_alloc(ctx, sizeof(double2) * ((size_t)m * 2 + (size_t)p.n), (void **)&bf);
    double2 *chirp = bf ? bf + 2 * m : NULL;
    if (!sb) {
      nxc_fft_filter_kernel<<<nxc_fft_grid(ctx, m), 256, 0, ctx->stream>>>(bf, p.n, m, p.sign);
      nxc_fft_chirp_kernel<<<nxc_fft_grid(ctx, p.n), 256, 0, ctx->stream>>>(chirp, p.n, p.sign);
      ctx->launches += 2;
      sb = nxc_fft_core(ctx, bf, bf + m, 1, m, -1);
    }
    NxcFftGather gw;
    NxcFftScatter sw;
    nxc_fft_contig(work, n_lines, m, &gw, &sw);
    // the lines kernel takes BOTH line bases from the gather's descriptor: the tensor's line numbering with
    // the work buffer's contiguous strides on the other side
    NxcFftLines to_work = ln, from_work = ln;
    {
      int64_t st = m;
      for (int i = ln.n - 1; i >= 0; i--) { to_work.dst_stride[i] = st; from_work.src_stride[i] = st; st *= ln.shape[i]; }
    }
    if (!sb) {
      NxcFftGather g;
      memset(&g, 0, sizeof g);
      g.lines = to_work;
      g.src = (const char *)src->data + src->offset * nxc_fft_esize(p.src_kind);
      g.n_lines = n_lines; g.n = p.n; g.n_src = p.n_src; g.m = m;
      g.stride = src->strides[p.axis];
      g.src_kind = p.src_kind; g.sign = p.sign; g.bluestein = 1; g.hermitian = p.hermitian; g.chirp = chirp;
      // the work lines take the source tensor's line numbering: same count, contiguous
      sb = nxc_fft_pow2(ctx, g, sw, n_lines, m, -1, mid);
    }
    if (!sb) {
      NxcFftScatter sc;
      memset(&sc, 0, sizeof sc);
      sc.lines = ln;
      sc.dst = (char *)dst->data + dst->offset * nxc_fft_esize(p.dst_kind);
      sc.n_lines = n_lines; sc.n = p.n; sc.n_out = p.n_out; sc.m = m;
      sc.stride = dst->strides[p.axis];
      sc.dst_kind = p.dst_kind; sc.sign = p.sign; sc.bluestein = 1; sc.chirp = chirp;
      gw.filter = bf;
      gw.lines = from_work;
      sb = nxc_fft_pow2(ctx, gw, sc, n_lines, m, +1, mid);
    }
    if (work) nxc_free(ctx, work);
    if (mid) nxc_free(ctx, mid);
    if (bf) nxc_free(ctx, bf);
    return sb;
  }
  const size_t wbytes = sizeof(double2) * (size_t)n_lines * (size_t)m;
  double2 *work = NULL, *tmp = NULL, *bf = NULL;
  nxc_status s = nxc_alloc(ctx, wbytes, (void **)&work);
  if (s) return s;
  if (m > NXC_FFT_SMEM_MAX) s = nxc_alloc(ctx, wbytes, (void **)&tmp);
  if (!s && !pow2) s = nxc_alloc(ctx, sizeof(double2) * (size_t)m * 2, (void **)&bf);
  if (!s) {
    NxcFftGather g;
    g.lines = ln;
    g.src = (const char *)src->data + src->offset * nxc_fft_esize(p.src_kind);
    g.work = work;
    g.n_lines = n_lines; g.n = p.n; g.n_src = p.n_src; g.m = m;
    g.stride = src->strides[p.axis];
    g.src_kind = p.src_kind; g.sign = p.sign; g.bluestein = !pow2; g.hermitian = p.hermitian; g.filter = NULL; g.chirp = NULL;
    nxc_fft_gather_kernel<<<nxc_fft_grid(ctx, n_lines * m), 256, 0, ctx->stream>>>(g);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "fft gather");
  }
  if (!s) {
    if (pow2) {
      s = nxc_fft_core(ctx, work, tmp, n_lines, m, p.sign);
    } else {
      nxc_fft_filter_kernel<<<nxc_fft_grid(ctx, m), 256, 0, ctx->stream>>>(bf, p.n, m, p.sign);
      ctx->launches++;
      s = nxc_fft_core(ctx, bf, bf + m, 1, m, -1);
      if (!s) s = nxc_fft_core(ctx, work, tmp, n_lines, m, -1);
      if (!s) {
        nxc_fft_pointwise_kernel<<<nxc_fft_grid(ctx, n_lines * m), 256, 0, ctx->stream>>>(work, bf, n_lines, m);
        ctx->launches++;
        s = nxc_fft_core(ctx, work, tmp, n_lines, m, +1);
      }
    }
  }
  if (!s) {
    NxcFftScatter g;
    g.lines = ln;
    g.dst = (char *)dst->data + dst->offset * nxc_fft_esize(p.dst_kind);
    g.work = work;
    g.n_lines = n_lines; g.n = p.n; g.n_out = p.n_out; g.m = m;
    g.stride = dst->strides[p.axis];
    g.dst_kind = p.dst_kind; g.sign = p.sign; g.bluestein = !pow2; g.chirp = NULL;
    nxc_fft_scatter_kernel<<<nxc_fft_grid(ctx, n_lines * p.n_out), 256, 0, ctx->stream>>>(g);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) s = nxc_cuda_fail(ctx, cudaGetLastError(), "fft scatter");
  }
  nxc_free(ctx, work);
  if (tmp) nxc_free(ctx, tmp);
  if (bf) nxc_free(ctx, bf);
  return s;
}

static int nxc_fft_kind(int dt) {
  switch (dt) {
    case NXC_C32: return NXC_FFT_SRC_C32;
    case NXC_C64: return NXC_FFT_SRC_C64;
    case NXC_F32: return NXC_FFT_SRC_F32;
    case NXC_F64: return NXC_FFT_SRC_F64;
    default: return -1;
  }
}

static nxc_status nxc_fft_fail(nxc_ctx *ctx, nxc_status s) {
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}

// fft / ifft (reference: nx_c_fft_run, nx_c_fft.c:940-955; dtype gate nx_c_fft.c:1166)
extern "C" nxc_status nxc_fft(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const int *axes, int n_axes,
                              int inverse) {
  NXC_TRACE(ctx, "nxc_fft");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) return nxc_fft_fail(ctx, s);
  if (in->dtype != NXC_C32 && in->dtype != NXC_C64) return nxc_fft_fail(ctx, NXC_ERR_BAD_KIND);
  if (out->dtype != in->dtype || out->ndim != in->ndim) return nxc_fft_fail(ctx, NXC_ERR_SHAPE);
  const int kind = nxc_fft_kind(in->dtype);
  for (int ai = 0; ai < n_axes; ai++) {
    const int axis = axes[ai];
    if (axis < 0 || axis >= in->ndim) return nxc_fft_fail(ctx, NXC_ERR_AXIS);
    NxcFftPass p;
    p.src = ai == 0 ? in : out;
    p.dst = out;
    p.src_kind = p.dst_kind = kind;
    p.axis = axis;
    p.n = p.n_src = p.n_out = out->shape[axis];
    p.sign = inverse ? 1 : -1;
    p.hermitian = 0;
    if ((s = nxc_fft_pass(ctx, p))) return nxc_fft_fail(ctx, s);
  }
  return NXC_OK;
}

// rfft (reference: nx_c_rfft_run, nx_c_fft.c:958-983)
extern "C" nxc_status nxc_rfft(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const int *axes, int n_axes) {
  NXC_TRACE(ctx, "nxc_rfft");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) return nxc_fft_fail(ctx, s);
  if ((in->dtype != NXC_F32 && in->dtype != NXC_F64) || (out->dtype != NXC_C32 && out->dtype != NXC_C64))
    return nxc_fft_fail(ctx, NXC_ERR_BAD_KIND);
  if (n_axes == 0) return NXC_OK;
  if (out->ndim != in->ndim) return nxc_fft_fail(ctx, NXC_ERR_SHAPE);
  const int last = axes[n_axes - 1];
  if (last < 0 || last >= in->ndim) return nxc_fft_fail(ctx, NXC_ERR_AXIS);
  const int64_t n = in->shape[last], half = n / 2 + 1;
  if (out->shape[last] != half) return nxc_fft_fail(ctx, NXC_ERR_SHAPE);
  NxcFftPass p;
  p.src = in; p.dst = out;
  p.src_kind = nxc_fft_kind(in->dtype); p.dst_kind = nxc_fft_kind(out->dtype);
  p.axis = last; p.n = p.n_src = n; p.n_out = half; p.sign = -1; p.hermitian = 0;
  if ((s = nxc_fft_pass(ctx, p))) return nxc_fft_fail(ctx, s);
  for (int ai = 0; ai < n_axes - 1; ai++) {
    const int axis = axes[ai];
    if (axis < 0 || axis >= out->ndim) return nxc_fft_fail(ctx, NXC_ERR_AXIS);
    p.src = out; p.dst = out;
    p.src_kind = p.dst_kind = nxc_fft_kind(out->dtype);
    p.axis = axis; p.n = p.n_src = p.n_out = out->shape[axis];
    if ((s = nxc_fft_pass(ctx, p))) return nxc_fft_fail(ctx, s);
  }
  return NXC_OK;
}

// irfft (reference: nx_c_irfft_run, nx_c_fft.c:1029-1143). s_last <= 0 infers 2*(half-1).
extern "C" nxc_status nxc_irfft(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const int *axes, int n_axes,
                                int64_t s_last) {
  NXC_TRACE(ctx, "nxc_irfft");
  nxc_status st;
  if ((st = nxc_check_tensor(in)) || (st = nxc_check_tensor(out))) return nxc_fft_fail(ctx, st);
  if ((in->dtype != NXC_C32 && in->dtype != NXC_C64) || (out->dtype != NXC_F32 && out->dtype != NXC_F64))
    return nxc_fft_fail(ctx, NXC_ERR_BAD_KIND);
  if (n_axes == 0) return NXC_OK;
  if (out->ndim != in->ndim) return nxc_fft_fail(ctx, NXC_ERR_SHAPE);
  const int last = axes[n_axes - 1];
  if (last < 0 || last >= in->ndim) return nxc_fft_fail(ctx, NXC_ERR_AXIS);
  const int64_t in_half = in->shape[last];
  const int64_t s = s_last > 0 ? s_last : 2 * (in_half - 1);
  if (out->shape[last] != s) return nxc_fft_fail(ctx, NXC_ERR_SHAPE);
  const int64_t needed_half = s / 2 + 1;
  const int64_t half = in_half < needed_half ? in_half : needed_half;
  for (int ai = 0; ai < n_axes - 1; ai++)
    if (axes[ai] < 0 || axes[ai] >= in->ndim) return nxc_fft_fail(ctx, NXC_ERR_AXIS);

  // the other transformed axes run in a contiguous c64 temporary of in's shape
  const nxc_tensor *spec = in;
  nxc_tensor tmp;
  void *tdata = NULL;
  int spec_kind = nxc_fft_kind(in->dtype);
  if (n_axes > 1) {
    tmp = *in;
    int64_t nelem = 1;
    for (int d = in->ndim - 1; d >= 0; d--) { tmp.strides[d] = nelem; nelem *= in->shape[d]; }
    tmp.offset = 0;
    tmp.dtype = NXC_C64;
    if ((st = nxc_alloc(ctx, sizeof(double2) * (size_t)(nelem ? nelem : 1), &tdata))) return nxc_fft_fail(ctx, st);
    tmp.data = tdata;
    for (int ai = 0; ai < n_axes - 1 && !st; ai++) {
      NxcFftPass p;
      p.src = ai == 0 ? in : &tmp;
      p.dst = &tmp;
      p.src_kind = ai == 0 ? nxc_fft_kind(in->dtype) : NXC_FFT_SRC_C64;
      p.dst_kind = NXC_FFT_SRC_C64;
      p.axis = axes[ai];
      p.n = p.n_src = p.n_out = in->shape[axes[ai]];
      p.sign = 1;
      p.hermitian = 0;
      st = nxc_fft_pass(ctx, p);
    }
    spec = &tmp;
    spec_kind = NXC_FFT_SRC_C64;
  }
  if (!st && s > 0) {
    NxcFftPass p;
    p.src = spec; p.dst = out;
    p.src_kind = spec_kind; p.dst_kind = nxc_fft_kind(out->dtype);
    p.axis = last; p.n = s; p.n_src = half; p.n_out = s; p.sign = 1; p.hermitian = 1;
    st = nxc_fft_pass(ctx, p);
  }
  if (tdata) nxc_free(ctx, tdata);
  return nxc_fft_fail(ctx, st);
}
