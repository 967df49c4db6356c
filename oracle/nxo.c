/* oracle/nxo.c -- TEST INFRASTRUCTURE ONLY. A CPU restatement of the reference's
 * Nx backend hot path (raven-ml/raven, packages/nx/lib/backend_c), written as a
 * small interpreter over type-erased scalars so that every rule is stated once
 * and is easy to audit against the reference file:line it follows. It is the
 * checker for the CUDA path; it is never shipped, never linked into
 * libnxcuda.so and never on the product path (only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it).
 *
 * PARITY PINNED: oracle/libnxo.so is compared, op by op over the reference
 * contract suite's pools and layout matrix, against the reference's own C
 * compiled unmodified (oracle/_ref/libnxref.so) by tests/test_oracle_pinning.py,
 * and against golden vectors generated from that binary (tests/golden/).
 *
 * What follows the reference, and where:
 *   dtype table, compute types, wrap/saturate policy   nx_c.h:130-168, 229-257, 345-363
 *   f16 / bf16 / fp8 converters                         buffer/nx_buffer_stubs.h:73-302
 *   unary / binary / compare / where expressions        nx_c_map.c:305-488, 499-740, 747-839
 *   cast matrix policy                                  nx_c_map.c:182-218
 *   reduce / argreduce / scan combines                  nx_c_fold.c:63-101, 110-198
 *   funnel + driver checks and statuses                 nx_c_engine.c:831-869, 1052-1105, 1214-1254, 1392-1420
 *   matmul shape rules, accumulate-in-compute-type      nx_c_matmul.c:874-936, 363-454
 *   pad / cat / gather / scatter                        nx_c_move.c:229-569
 *   threefry2x32-20                                     nx_c_random.c:44-61
 *
 * How elements are walked is NOT the reference's (no coalescing, no thread
 * pool, no blocked GEMM): a plain odometer, one element at a time. Float sums use
 * the reference's 16-accumulator balanced tree on contiguous runs
 * (nx_c_fold.c:139-159) but never its streaming path, so float sums agree with
 * the reference to reassociation error only (the contract's own tolerance);
 * everything else is bit-for-bit.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NXO_MAX_NDIM 32

typedef struct {
  void *data;
  int32_t dtype;
  int32_t ndim;
  int64_t shape[NXO_MAX_NDIM];
  int64_t strides[NXO_MAX_NDIM];
  int64_t offset;
} nxo_tensor;

typedef const char *nxo_status;

/* status texts: nx_c.h:381-388, nx_c_engine.h:38-50, nx_c_matmul.c:867-868 */
#define E_NDIM "ndim exceeds NX_C_MAX_NDIM"
#define E_BAD_KIND "unsupported bigarray kind"
#define E_UNSUPPORTED "dtype not supported for this operation"
#define E_PACKED "packed dtype not supported for this operation"
#define E_SHAPE "shape mismatch"
#define E_EMPTY_REDUCE "reduction over empty axis has no identity"
#define E_ARGCAP "argreduce axis length exceeds INT32_MAX"
#define E_AXES "reduce axes must be strictly increasing and in range"
#define E_OUT_RANK "output rank inconsistent with the operation"
#define E_AXIS "axis out of range"
#define E_ALIASED "output has a broadcast (zero) stride"
#define E_MM_DTYPE "matmul operands must share one dtype"
#define E_BAD_OP "unknown operation code"
#define E_INDEX_OOB "index out of bounds for the gathered/scattered axis" /* nx_c_move.c:49 */
#define E_THREEFRY_SHAPE "threefry: last axis must have extent 2"             /* nx_c_random.c:32 */

enum { F16, F32, F64, BF16, F8E4M3, F8E5M2, I4, U4, I8, U8, I16, U16, I32, U32, I64, U64, C32, C64, BOOL_, NDT };
/* compute kinds = the reference's compute types (nx_c.h:130-168) */
enum { K_F32, K_F64, K_I64, K_U64, K_C32, K_C64, K_BOOL, K_NONE };
enum { CAT_SINT = 1, CAT_UINT = 2, CAT_FLOAT = 4, CAT_COMPLEX = 8, CAT_BOOL = 16, CAT_PACKED = 32 };

static const int DT_KIND[NDT] = {K_F32, K_F32, K_F64, K_F32, K_F32, K_F32, K_NONE, K_NONE, K_I64, K_I64,
                                 K_I64, K_I64, K_I64, K_U64, K_I64, K_U64, K_C32, K_C64, K_BOOL};
static const int DT_SIZE[NDT] = {2, 4, 8, 2, 1, 1, 0, 0, 1, 1, 2, 2, 4, 4, 8, 8, 8, 16, 1};
static const int DT_CAT[NDT] = {CAT_FLOAT, CAT_FLOAT, CAT_FLOAT, CAT_FLOAT, CAT_FLOAT, CAT_FLOAT,
                                CAT_SINT | CAT_PACKED, CAT_UINT | CAT_PACKED, CAT_SINT, CAT_UINT, CAT_SINT,
                                CAT_UINT, CAT_SINT, CAT_UINT, CAT_SINT, CAT_UINT, CAT_COMPLEX, CAT_COMPLEX,
                                CAT_BOOL};

int64_t nxo_elem_size(int dt) { return (dt >= 0 && dt < NDT) ? DT_SIZE[dt] : 0; }

typedef union {
  float f;
  double d;
  int64_t i;
  uint64_t u;
  float _Complex c32;
  double _Complex c64;
  uint8_t b;
} val;

/* ---- storage converters (buffer/nx_buffer_stubs.h:73-302) --------------------- */
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static uint16_t to_bf16(float f) {
  uint32_t b = f2u(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((b >> 16) | 0x40u); /* quiet, keep sign */
  return (uint16_t)((b + (((b >> 16) & 1u) + 0x7FFFu)) >> 16);               /* RNE */
}
static float from_bf16(uint16_t h) { return u2f((uint32_t)h << 16); }

static uint16_t to_f16(float f) {
  uint32_t b = f2u(f), sgn = (b >> 16) & 0x8000u, e = b & 0x7F800000u, m = b & 0x7FFFFFu;
  if (e >= 0x47800000u) { /* overflow, inf, NaN (payload kept, never turns into inf) */
    if (e == 0x7F800000u && m) {
      uint16_t r = (uint16_t)(0x7C00u + (m >> 13));
      if (r == 0x7C00u) r++;
      return (uint16_t)(sgn + r);
    }
    return (uint16_t)(sgn + 0x7C00u);
  }
  if (e <= 0x38000000u) { /* subnormal half or zero */
    if (e < 0x33000000u) return (uint16_t)sgn;
    uint32_t ex = e >> 23, s = (m + 0x800000u) >> (113 - ex);
    if (((s & 0x3FFFu) != 0x1000u) || (b & 0x7FFu)) s += 0x1000u; /* RNE with sticky low bits */
    return (uint16_t)(sgn + (s >> 13));
  }
  if ((m & 0x3FFFu) != 0x1000u) m += 0x1000u; /* RNE; a carry bumps the exponent */
  return (uint16_t)(sgn + ((e - 0x38000000u) >> 13) + (m >> 13));
}
static float from_f16(uint16_t h) {
  uint32_t s = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  if (e == 0x1F) return u2f(s | 0x7F800000u | (m ? ((m << 13) | 0x400000u) : 0));
  if (e == 0) {
    if (!m) return u2f(s);
    int ex = 1;
    while (!(m & 0x400u)) { m <<= 1; ex--; }
    return u2f(s | ((uint32_t)(ex + 112) << 23) | ((m & 0x3FFu) << 13));
  }
  return u2f(s | ((e + 112) << 23) | (m << 13));
}
/* EB/MB small floats, RNE, subnormals; e4m3 is the "fn" flavour (no inf, overflow -> NaN) */
static uint8_t to_fp8(float f, int mb, int bias, uint32_t maxbits, uint32_t ovf, uint32_t infb) {
  if (isnan(f)) return 0x7F;
  uint32_t b = f2u(f), sign = (b >> 31) << 7;
  if (isinf(f)) return (uint8_t)(sign | infb);
  int ex = (int)((b >> 23) & 0xFF) - 127, emin = 1 - bias;
  if (ex >= emin) {
    int sh = 23 - mb;
    uint32_t sig = b & 0x7FFFFFu, q = sig >> sh, rem = sig & ((1u << sh) - 1u), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t bits = ((uint32_t)(ex + bias) << mb) + q;
    return (uint8_t)(bits >= maxbits ? (sign | ovf) : (sign | bits));
  }
  int shift = (23 - mb) + (emin - ex);
  if (shift > 24) return (uint8_t)sign;
  uint32_t sig = (b & 0x7FFFFFu) | 0x800000u, q = sig >> shift, rem = sig & ((1u << shift) - 1u),
           half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) q++;
  return (uint8_t)(sign | q);
}
static float from_e4m3(uint8_t v) {
  uint32_t e = (v >> 3) & 0xF, m = v & 7;
  if (e == 0xF && m == 7) return NAN;
  float r = e ? ldexpf(1.0f + (float)m / 8.0f, (int)e - 7) : ldexpf((float)m, -9);
  return (v & 0x80) ? -r : r;
}
static float from_e5m2(uint8_t v) {
  uint32_t e = (v >> 2) & 0x1F, m = v & 3;
  if (e == 0x1F) return m ? NAN : ((v & 0x80) ? -INFINITY : INFINITY);
  float r = e ? ldexpf(1.0f + (float)m / 4.0f, (int)e - 15) : ldexpf((float)m / 4.0f, -14);
  return (v & 0x80) ? -r : r;
}

/* ---- load / store in the compute type (nx_c.h:219-227) -------------------------- */
static val ld(int dt, const char *p) {
  val v;
  memset(&v, 0, sizeof v);
  switch (dt) {
    case F16: v.f = from_f16(*(const uint16_t *)p); break;
    case F32: v.f = *(const float *)p; break;
    case F64: v.d = *(const double *)p; break;
    case BF16: v.f = from_bf16(*(const uint16_t *)p); break;
    case F8E4M3: v.f = from_e4m3(*(const uint8_t *)p); break;
    case F8E5M2: v.f = from_e5m2(*(const uint8_t *)p); break;
    case I8: v.i = *(const int8_t *)p; break;
    case U8: v.i = *(const uint8_t *)p; break;
    case I16: v.i = *(const int16_t *)p; break;
    case U16: v.i = *(const uint16_t *)p; break;
    case I32: v.i = *(const int32_t *)p; break;
    case U32: v.u = *(const uint32_t *)p; break;
    case I64: v.i = *(const int64_t *)p; break;
    case U64: v.u = *(const uint64_t *)p; break;
    case C32: v.c32 = *(const float _Complex *)p; break;
    case C64: v.c64 = *(const double _Complex *)p; break;
    case BOOL_: v.b = (*(const uint8_t *)p != 0); break;
  }
  return v;
}
static void st(int dt, char *p, val v) {
  switch (dt) {
    case F16: *(uint16_t *)p = to_f16(v.f); break;
    case F32: *(float *)p = v.f; break;
    case F64: *(double *)p = v.d; break;
    case BF16: *(uint16_t *)p = to_bf16(v.f); break;
    case F8E4M3: *(uint8_t *)p = to_fp8(v.f, 3, 7, 0x7F, 0x7F, 0x7F); break;
    case F8E5M2: *(uint8_t *)p = to_fp8(v.f, 2, 15, 0x7C, 0x7C, 0x7C); break;
    case I8: *(int8_t *)p = (int8_t)v.i; break; /* integer stores wrap (nx_c.h:349-350) */
    case U8: *(uint8_t *)p = (uint8_t)v.i; break;
    case I16: *(int16_t *)p = (int16_t)v.i; break;
    case U16: *(uint16_t *)p = (uint16_t)v.i; break;
    case I32: *(int32_t *)p = (int32_t)v.i; break;
    case U32: *(uint32_t *)p = (uint32_t)v.u; break;
    case I64: *(int64_t *)p = v.i; break;
    case U64: *(uint64_t *)p = v.u; break;
    case C32: *(float _Complex *)p = v.c32; break;
    case C64: *(double _Complex *)p = v.c64; break;
    case BOOL_: *(uint8_t *)p = (v.b != 0); break;
  }
}

/* ---- odometer over a shape with several operands -------------------------------- */
typedef struct {
  int ndim, nop;
  int64_t shape[NXO_MAX_NDIM], coord[NXO_MAX_NDIM];
  int64_t bstride[6][NXO_MAX_NDIM];
  char *ptr[6];
  int64_t total;
} odo;

static void odo_init(odo *o, int ndim, const int64_t *shape) {
  o->ndim = ndim;
  o->nop = 0;
  o->total = 1;
  for (int i = 0; i < ndim; i++) { o->shape[i] = shape[i]; o->coord[i] = 0; o->total *= shape[i]; }
}
static void odo_add(odo *o, const nxo_tensor *t, const int *dims /* NULL = identity */) {
  int k = o->nop++;
  int64_t es = DT_SIZE[t->dtype];
  o->ptr[k] = (char *)t->data + t->offset * es;
  for (int i = 0; i < o->ndim; i++) o->bstride[k][i] = t->strides[dims ? dims[i] : i] * es;
}
static void odo_next(odo *o) {
  for (int d = o->ndim - 1; d >= 0; d--) {
    if (++o->coord[d] < o->shape[d]) {
      for (int k = 0; k < o->nop; k++) o->ptr[k] += o->bstride[k][d];
      return;
    }
    o->coord[d] = 0;
    for (int k = 0; k < o->nop; k++) o->ptr[k] -= (o->shape[d] - 1) * o->bstride[k][d];
  }
}

static nxo_status chk(const nxo_tensor *t) {
  if (t->ndim < 0 || t->ndim > NXO_MAX_NDIM) return E_NDIM;
  if (t->dtype < 0 || t->dtype >= NDT) return E_BAD_KIND;
  return NULL;
}
static nxo_status out_aliased(const nxo_tensor *out) {
  int64_t total = 1;
  for (int i = 0; i < out->ndim; i++) total *= out->shape[i];
  if (total == 0) return NULL;
  for (int i = 0; i < out->ndim; i++)
    if (out->shape[i] > 1 && out->strides[i] == 0) return E_ALIASED;
  return NULL;
}

/* ---- integer power (nx_c_map.c:147-165) ------------------------------------------- */
static int64_t ipow_s(int64_t base, int64_t e) {
  if (e < 0) return base == 1 ? 1 : base == -1 ? ((e & 1) ? -1 : 1) : 0;
  uint64_t b = (uint64_t)base, r = 1;
  while (e > 0) { if (e & 1) r *= b; e >>= 1; if (e) b *= b; }
  return (int64_t)r;
}
static uint64_t ipow_u(uint64_t b, uint64_t e) {
  uint64_t r = 1;
  while (e > 0) { if (e & 1) r *= b; e >>= 1; if (e) b *= b; }
  return r;
}

/* ==== unary ========================================================================= */
enum { NEG, RECIP, ABS, SIGN, SQRT, EXP, LOG, SIN, COS, TAN, ASIN, ACOS, ATAN, SINH, COSH, TANH, TRUNC, CEIL,
       FLOOR, ROUND, ERF, N_UN };

static int un_mask(int op) { /* table rows: nx_c_map.c:224-294, 368-488 */
  switch (op) {
    case NEG: case RECIP: case ABS: case SIGN: return CAT_SINT | CAT_UINT | CAT_FLOAT | CAT_COMPLEX;
    case ERF: return CAT_FLOAT;
    case TRUNC: case CEIL: case FLOOR: case ROUND: return CAT_SINT | CAT_UINT | CAT_FLOAT;
    default: return CAT_FLOAT | CAT_COMPLEX;
  }
}

#define UN_FLOAT(T, SFX, x)                                                              \
  switch (op) {                                                                          \
    case NEG: return -(x);                                                               \
    case RECIP: return (T)1 / (x);                                                       \
    case ABS: return fabs##SFX(x);                                                       \
    case SIGN: return isnan(x) ? (x) : (T)(((x) > 0) - ((x) < 0));                       \
    case SQRT: return sqrt##SFX(x); case EXP: return exp##SFX(x); case LOG: return log##SFX(x); \
    case SIN: return sin##SFX(x); case COS: return cos##SFX(x); case TAN: return tan##SFX(x);   \
    case ASIN: return asin##SFX(x); case ACOS: return acos##SFX(x); case ATAN: return atan##SFX(x); \
    case SINH: return sinh##SFX(x); case COSH: return cosh##SFX(x); case TANH: return tanh##SFX(x); \
    case TRUNC: return trunc##SFX(x); case CEIL: return ceil##SFX(x); case FLOOR: return floor##SFX(x); \
    case ROUND: return round##SFX(x); case ERF: return erf##SFX(x);                      \
  }
static float un_f32(int op, float x) { UN_FLOAT(float, f, x) return x; }
static double un_f64(int op, double x) { UN_FLOAT(double, , x) return x; }

#define UN_CPLX(T, R, SFX, x)                                                            \
  switch (op) {                                                                          \
    case NEG: return -(x);                                                               \
    case RECIP: return (T)1 / (x);                                                       \
    case ABS: return (T)cabs##SFX(x);                                                    \
    case SIGN: return cabs##SFX(x) == 0 ? (T)0 : (x) / cabs##SFX(x);                     \
    case SQRT: return csqrt##SFX(x); case EXP: return cexp##SFX(x); case LOG: return clog##SFX(x); \
    case SIN: return csin##SFX(x); case COS: return ccos##SFX(x); case TAN: return ctan##SFX(x);   \
    case ASIN: return casin##SFX(x); case ACOS: return cacos##SFX(x); case ATAN: return catan##SFX(x); \
    case SINH: return csinh##SFX(x); case COSH: return ccosh##SFX(x); case TANH: return ctanh##SFX(x); \
  }
static float _Complex un_c32(int op, float _Complex x) { UN_CPLX(float _Complex, float, f, x) return x; }
static double _Complex un_c64(int op, double _Complex x) { UN_CPLX(double _Complex, double, , x) return x; }

static val un_apply(int op, int dt, val x) {
  val r = x;
  switch (DT_KIND[dt]) {
    case K_F32: r.f = un_f32(op, x.f); break;
    case K_F64: r.d = un_f64(op, x.d); break;
    case K_C32: r.c32 = un_c32(op, x.c32); break;
    case K_C64: r.c64 = un_c64(op, x.c64); break;
    case K_I64: /* signed ints and u8/u16: nx_c_map.c:305-366; negate in the unsigned width */
      if (DT_CAT[dt] & CAT_SINT) {
        if (op == NEG) r.i = (int64_t)(-(uint64_t)x.i);
        else if (op == RECIP) r.i = x.i == 0 ? 0 : 1 / x.i;
        else if (op == ABS) r.i = x.i < 0 ? (int64_t)(-(uint64_t)x.i) : x.i;
        else if (op == SIGN) r.i = (x.i > 0) - (x.i < 0);
      } else {
        if (op == NEG) r.i = -x.i;
        else if (op == RECIP) r.i = x.i == 0 ? 0 : 1 / x.i;
        else if (op == SIGN) r.i = (x.i != 0);
      }
      break;
    case K_U64:
      if (op == NEG) r.u = -x.u;
      else if (op == RECIP) r.u = x.u == 0 ? 0 : 1 / x.u;
      else if (op == SIGN) r.u = (x.u != 0);
      break;
  }
  return r;
}

nxo_status nxo_map1(int op, const nxo_tensor *out, const nxo_tensor *a) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(a))) return s;
  if ((DT_CAT[out->dtype] | DT_CAT[a->dtype]) & CAT_PACKED) return E_PACKED;
  if (op < 0 || op >= N_UN) return E_BAD_OP;
  int dt = out->dtype;
  if (!(un_mask(op) & DT_CAT[dt])) return E_UNSUPPORTED;
  if ((s = out_aliased(out))) return s;
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, a, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) st(dt, o.ptr[0], un_apply(op, dt, ld(dt, o.ptr[1])));
  return NULL;
}

/* ==== binary ======================================================================== */
enum { ADD, SUB, MUL, IDIV, FDIV, MOD, MAX, MIN, POW, ATAN2, XOR, OR, AND, SHL, SHR, N_BIN };

static int bin_mask(int op) { /* nx_c_map.c:499-740 */
  switch (op) {
    case ADD: case SUB: case MUL: case POW: return CAT_SINT | CAT_UINT | CAT_FLOAT | CAT_COMPLEX;
    case IDIV: case MOD: return CAT_SINT | CAT_UINT | CAT_FLOAT;
    case FDIV: return CAT_FLOAT | CAT_COMPLEX;
    case MAX: case MIN: return CAT_SINT | CAT_UINT | CAT_FLOAT | CAT_BOOL;
    case ATAN2: return CAT_FLOAT;
    case XOR: case OR: case AND: return CAT_SINT | CAT_UINT | CAT_BOOL;
    case SHL: case SHR: return CAT_SINT | CAT_UINT;
  }
  return 0;
}

#define BIN_FLOAT(T, SFX, a, b)                                                        \
  switch (op) {                                                                        \
    case ADD: return (a) + (b); case SUB: return (a) - (b); case MUL: return (a) * (b); \
    case IDIV: return trunc##SFX((a) / (b));                                           \
    case FDIV: return (a) / (b);                                                       \
    case MOD: return fmod##SFX(a, b);                                                  \
    case MAX: return (isnan(a) || isnan(b)) ? (T)NAN : ((a) > (b) ? (a) : (b));        \
    case MIN: return (isnan(a) || isnan(b)) ? (T)NAN : ((a) < (b) ? (a) : (b));        \
    case POW: return pow##SFX(a, b);                                                   \
    case ATAN2: return atan2##SFX(a, b);                                               \
  }
static float bin_f32(int op, float a, float b) { BIN_FLOAT(float, f, a, b) return a; }
static double bin_f64(int op, double a, double b) { BIN_FLOAT(double, , a, b) return a; }

static val bin_apply(int op, int dt, val a, val b) {
  val r = a;
  const int bits = DT_SIZE[dt] * 8;
  switch (DT_KIND[dt]) {
    case K_F32: r.f = bin_f32(op, a.f, b.f); break;
    case K_F64: r.d = bin_f64(op, a.d, b.d); break;
    case K_C32:
      r.c32 = op == ADD ? a.c32 + b.c32 : op == SUB ? a.c32 - b.c32 : op == MUL ? a.c32 * b.c32
              : op == FDIV ? a.c32 / b.c32 : cpowf(a.c32, b.c32);
      break;
    case K_C64:
      r.c64 = op == ADD ? a.c64 + b.c64 : op == SUB ? a.c64 - b.c64 : op == MUL ? a.c64 * b.c64
              : op == FDIV ? a.c64 / b.c64 : cpow(a.c64, b.c64);
      break;
    case K_BOOL:
      r.b = op == MAX ? (a.b > b.b ? a.b : b.b) : op == MIN ? (a.b < b.b ? a.b : b.b)
            : op == XOR ? (a.b ^ b.b) : op == OR ? (a.b | b.b) : (a.b & b.b);
      break;
    case K_I64: {
      int64_t x = a.i, y = b.i;
      int sgn = (DT_CAT[dt] & CAT_SINT) != 0;
      switch (op) {
        case ADD: r.i = (int64_t)((uint64_t)x + (uint64_t)y); break;
        case SUB: r.i = (int64_t)((uint64_t)x - (uint64_t)y); break;
        case MUL: r.i = (int64_t)((uint64_t)x * (uint64_t)y); break;
        case IDIV: r.i = y == 0 ? 0 : (sgn && y == -1) ? (int64_t)(-(uint64_t)x) : x / y; break;
        case MOD: r.i = y == 0 ? 0 : (sgn && y == -1) ? 0 : x % y; break;
        case MAX: r.i = x > y ? x : y; break;
        case MIN: r.i = x < y ? x : y; break;
        case POW: r.i = sgn ? ipow_s(x, y) : (int64_t)ipow_u((uint64_t)x, (uint64_t)y); break;
        case XOR: r.i = x ^ y; break;
        case OR: r.i = x | y; break;
        case AND: r.i = x & y; break;
        case SHL: r.i = ((sgn && y < 0) || y >= bits) ? 0 : (int64_t)((uint64_t)x << y); break;
        case SHR: r.i = ((sgn && y < 0) || y >= bits) ? 0 : (x >> y); break;
      }
    } break;
    case K_U64: {
      uint64_t x = a.u, y = b.u;
      switch (op) {
        case ADD: r.u = x + y; break; case SUB: r.u = x - y; break; case MUL: r.u = x * y; break;
        case IDIV: r.u = y == 0 ? 0 : x / y; break;
        case MOD: r.u = y == 0 ? 0 : x % y; break;
        case MAX: r.u = x > y ? x : y; break;
        case MIN: r.u = x < y ? x : y; break;
        case POW: r.u = ipow_u(x, y); break;
        case XOR: r.u = x ^ y; break; case OR: r.u = x | y; break; case AND: r.u = x & y; break;
        case SHL: r.u = y >= (uint64_t)bits ? 0 : x << y; break;
        case SHR: r.u = y >= (uint64_t)bits ? 0 : x >> y; break;
      }
    } break;
  }
  return r;
}

nxo_status nxo_map2(int op, const nxo_tensor *out, const nxo_tensor *a, const nxo_tensor *b) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(a)) || (s = chk(b))) return s;
  if ((DT_CAT[out->dtype] | DT_CAT[a->dtype] | DT_CAT[b->dtype]) & CAT_PACKED) return E_PACKED;
  if (op < 0 || op >= N_BIN) return E_BAD_OP;
  int dt = out->dtype;
  if (!(bin_mask(op) & DT_CAT[dt])) return E_UNSUPPORTED;
  if ((s = out_aliased(out))) return s;
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, a, NULL);
  odo_add(&o, b, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o))
    st(dt, o.ptr[0], bin_apply(op, dt, ld(dt, o.ptr[1]), ld(dt, o.ptr[2])));
  return NULL;
}

/* ==== comparisons (nx_c_map.c:747-803, dispatched on the INPUT dtype) ================== */
enum { CMPEQ, CMPNE, CMPLT, CMPLE };
static int cmp_apply(int op, int dt, val a, val b) {
  switch (DT_KIND[dt]) {
#define C4(x, y) (op == CMPEQ ? (x) == (y) : op == CMPNE ? (x) != (y) : op == CMPLT ? (x) < (y) : (x) <= (y))
    case K_F32: return C4(a.f, b.f);
    case K_F64: return C4(a.d, b.d);
    case K_I64: return C4(a.i, b.i);
    case K_U64: return C4(a.u, b.u);
    case K_BOOL: return C4(a.b, b.b);
#undef C4
    case K_C32: return op == CMPEQ ? a.c32 == b.c32 : a.c32 != b.c32;
    case K_C64: return op == CMPEQ ? a.c64 == b.c64 : a.c64 != b.c64;
  }
  return 0;
}
nxo_status nxo_cmp(int op, const nxo_tensor *out, const nxo_tensor *a, const nxo_tensor *b) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(a)) || (s = chk(b))) return s;
  int dt = a->dtype;
  if ((DT_CAT[dt] | DT_CAT[out->dtype]) & CAT_PACKED) return E_PACKED;
  if (op < 0 || op > CMPLE) return E_BAD_OP;
  if ((op == CMPLT || op == CMPLE) && (DT_CAT[dt] & CAT_COMPLEX)) return E_UNSUPPORTED;
  if ((s = out_aliased(out))) return s;
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, a, NULL);
  odo_add(&o, b, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o))
    *(uint8_t *)o.ptr[0] = (uint8_t)cmp_apply(op, dt, ld(dt, o.ptr[1]), ld(dt, o.ptr[2]));
  return NULL;
}

/* ==== where: bit-exact select on the storage type (nx_c_map.c:809-839) ================= */
nxo_status nxo_where(const nxo_tensor *out, const nxo_tensor *c, const nxo_tensor *a, const nxo_tensor *b) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(c)) || (s = chk(a)) || (s = chk(b))) return s;
  if (DT_CAT[out->dtype] & CAT_PACKED) return E_PACKED;
  if ((s = out_aliased(out))) return s;
  int es = DT_SIZE[out->dtype];
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, c, NULL);
  odo_add(&o, a, NULL);
  odo_add(&o, b, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o))
    memcpy(o.ptr[0], *(uint8_t *)o.ptr[1] ? o.ptr[2] : o.ptr[3], (size_t)es);
  return NULL;
}

/* ==== copy (nx_c_move.c:61-84) ========================================================== */
static nxo_status copy_packed(const nxo_tensor *out, const nxo_tensor *in);
nxo_status nxo_copy(const nxo_tensor *out, const nxo_tensor *a) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(a))) return s;
  if (DT_CAT[out->dtype] & CAT_PACKED) return copy_packed(out, a);
  if ((s = out_aliased(out))) return s;
  int es = DT_SIZE[out->dtype];
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, a, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) memcpy(o.ptr[0], o.ptr[1], (size_t)es);
  return NULL;
}

/* ==== cast (policy: nx_c_map.c:182-218; saturation: nx_c.h:234-250) ==================== */
static int64_t f2i_s(double v, int w) {
  double lim = ldexp(1.0, w - 1);
  if (isnan(v)) return 0;
  if (v <= -lim) return w == 64 ? INT64_MIN : -((int64_t)1 << (w - 1));
  if (v >= lim) return (int64_t)(((uint64_t)1 << (w - 1)) - 1);
  return (int64_t)v;
}
static uint64_t f2i_u(double v, int w) {
  if (isnan(v) || v <= 0.0) return 0;
  if (v >= ldexp(1.0, w)) return w == 64 ? ~(uint64_t)0 : (((uint64_t)1 << w) - 1);
  return (uint64_t)v;
}
static val cast_apply(int src, int dst, val v) {
  val r;
  memset(&r, 0, sizeof r);
  int sk = DT_KIND[src], dk = DT_KIND[dst];
  /* the source as a real double / as exact integers, whichever the rule needs */
  double re = 0, im = 0;
  switch (sk) {
    case K_F32: re = v.f; break;
    case K_F64: re = v.d; break;
    case K_I64: re = (double)v.i; break;
    case K_U64: re = (double)v.u; break;
    case K_C32: re = crealf(v.c32); im = cimagf(v.c32); break;
    case K_C64: re = creal(v.c64); im = cimag(v.c64); break;
    case K_BOOL: re = v.b; break;
  }
  int src_is_float = (sk == K_F32 || sk == K_F64 || sk == K_C32 || sk == K_C64);
  switch (dk) {
    case K_F32: /* (float)(real part); ints convert directly from their integer type */
      r.f = sk == K_I64 ? (float)v.i : sk == K_U64 ? (float)v.u : sk == K_F32 ? v.f : sk == K_C32 ? crealf(v.c32)
            : (float)re;
      break;
    case K_F64: r.d = sk == K_I64 ? (double)v.i : sk == K_U64 ? (double)v.u : re; break;
    case K_C32:
      r.c32 = sk == K_C32 ? v.c32 : sk == K_C64 ? (float _Complex)v.c64
              : sk == K_I64 ? (float _Complex)v.i : sk == K_U64 ? (float _Complex)v.u
              : sk == K_F32 ? (float _Complex)v.f : (float _Complex)re;
      break;
    case K_C64:
      r.c64 = sk == K_C64 ? v.c64 : sk == K_C32 ? (double _Complex)v.c32
              : sk == K_I64 ? (double _Complex)v.i : sk == K_U64 ? (double _Complex)v.u : (double _Complex)re;
      break;
    case K_BOOL: r.b = (sk == K_C32 || sk == K_C64) ? (re != 0 || im != 0) : sk == K_I64 ? (v.i != 0)
                       : sk == K_U64 ? (v.u != 0) : (re != 0); /* NaN -> true */
      break;
    case K_I64:
      if (src_is_float) {
        int w = DT_SIZE[dst] * 8;
        r.i = (DT_CAT[dst] & CAT_SINT) ? f2i_s(re, w) : (int64_t)f2i_u(re, w);
      } else {
        r.i = sk == K_U64 ? (int64_t)v.u : sk == K_BOOL ? v.b : v.i; /* wraps on store */
      }
      break;
    case K_U64:
      if (src_is_float) r.u = f2i_u(re, DT_SIZE[dst] * 8);
      else r.u = sk == K_U64 ? v.u : sk == K_BOOL ? v.b : (uint64_t)v.i;
      break;
  }
  return r;
}
static nxo_status cast_packed(const nxo_tensor *out, const nxo_tensor *in);
nxo_status nxo_cast(const nxo_tensor *out, const nxo_tensor *a) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(a))) return s;
  if ((DT_CAT[out->dtype] | DT_CAT[a->dtype]) & CAT_PACKED) return cast_packed(out, a);
  if ((s = out_aliased(out))) return s;
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, a, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o))
    st(out->dtype, o.ptr[0], cast_apply(a->dtype, out->dtype, ld(a->dtype, o.ptr[1])));
  return NULL;
}


/* ==== packed int4 / uint4 (storage-only): nx_c_map.c:1046-1177, nx_c_move.c:151-189 ===== */
static int packed_dense(const nxo_tensor *a) {
  int64_t expect = 1;
  for (int i = a->ndim - 1; i >= 0; i--) {
    if (a->shape[i] == 1) continue;
    if (a->strides[i] != expect) return 0;
    expect *= a->shape[i];
  }
  return 1;
}
static int get_nib(const uint8_t *b, int64_t i, int sgn) {
  uint8_t by = b[i >> 1];
  if (sgn) return (i & 1) ? ((int8_t)by >> 4) : ((int8_t)((by & 0x0F) << 4) >> 4);
  return (i & 1) ? (by >> 4) : (by & 0x0F);
}
static void put_nib(uint8_t *b, int64_t i, unsigned nib) {
  uint8_t *bp = &b[i >> 1];
  *bp = (i & 1) ? (uint8_t)((*bp & 0x0F) | ((nib & 0xF) << 4)) : (uint8_t)((*bp & 0xF0) | (nib & 0xF));
}
static int f2i4(double v, int sgn) {
  if (isnan(v)) return 0;
  if (sgn) { if (v <= -8.0) return -8; if (v >= 7.0) return 7; }
  else { if (v <= 0.0) return 0; if (v >= 15.0) return 15; }
  return (int)v;
}
static nxo_status cast_packed(const nxo_tensor *out, const nxo_tensor *in) {
  if (!packed_dense(out) || !packed_dense(in)) return E_PACKED;
  int64_t n = 1;
  for (int i = 0; i < out->ndim; i++) n *= out->shape[i];
  int src = in->dtype, dst = out->dtype;
  int sp = (DT_CAT[src] & CAT_PACKED) != 0, dp = (DT_CAT[dst] & CAT_PACKED) != 0;
  for (int64_t i = 0; i < n; i++) {
    if (sp && dp) {
      put_nib((uint8_t *)out->data, out->offset + i, (unsigned)get_nib((const uint8_t *)in->data, in->offset + i, 0));
    } else if (dp) {
      val v = ld(src, (const char *)in->data + (in->offset + i) * DT_SIZE[src]);
      int sgn = dst == I4, nib;
      switch (DT_KIND[src]) {
        case K_F32: nib = f2i4((double)v.f, sgn); break;
        case K_F64: nib = f2i4(v.d, sgn); break;
        case K_C32: nib = f2i4((double)crealf(v.c32), sgn); break;
        case K_C64: nib = f2i4(creal(v.c64), sgn); break;
        case K_U64: nib = (int)v.u; break;
        case K_BOOL: nib = v.b; break;
        default: nib = (int)v.i; break;
      }
      put_nib((uint8_t *)out->data, out->offset + i, (unsigned)nib);
    } else {
      int nv = get_nib((const uint8_t *)in->data, in->offset + i, src == I4);
      val v;
      memset(&v, 0, sizeof v);
      v.i = nv; /* the nibble as a small signed / unsigned integer, then the cast policy */
      st(dst, (char *)out->data + (out->offset + i) * DT_SIZE[dst], cast_apply(src == I4 ? I8 : U8, dst, v));
    }
  }
  return NULL;
}
static nxo_status copy_packed(const nxo_tensor *out, const nxo_tensor *in) {
  int64_t n = 1, m = 1;
  for (int i = 0; i < out->ndim; i++) n *= out->shape[i];
  for (int i = 0; i < in->ndim; i++) m *= in->shape[i];
  if (n != m || out->offset != 0 || in->offset != 0 || !packed_dense(out) || !packed_dense(in)) return E_PACKED;
  memcpy(out->data, in->data, (size_t)(n / 2));
  if (n & 1) { /* odd tail: merge only the low nibble, the neighbour's high nibble survives */
    uint8_t *d = (uint8_t *)out->data + n / 2;
    const uint8_t *sp = (const uint8_t *)in->data + n / 2;
    *d = (uint8_t)((*d & 0xF0) | (*sp & 0x0F));
  }
  return NULL;
}

/* ==== fold family ======================================================================= */
enum { R_SUM, R_PROD, R_MAX, R_MIN };

static val red_init(int op, int dt) { /* nx_c_fold.c:50-61, 261-314 */
  val v;
  memset(&v, 0, sizeof v);
  switch (DT_KIND[dt]) {
    case K_F32: v.f = op == R_SUM ? 0 : op == R_PROD ? 1 : op == R_MAX ? -INFINITY : INFINITY; break;
    case K_F64: v.d = op == R_SUM ? 0 : op == R_PROD ? 1 : op == R_MAX ? -INFINITY : INFINITY; break;
    case K_I64: v.i = op == R_SUM ? 0 : op == R_PROD ? 1 : op == R_MAX ? INT64_MIN : INT64_MAX; break;
    case K_U64: v.u = op == R_SUM ? 0 : op == R_PROD ? 1 : op == R_MAX ? 0 : UINT64_MAX; break;
    case K_C32: v.c32 = op == R_PROD ? 1 : 0; break;
    case K_C64: v.c64 = op == R_PROD ? 1 : 0; break;
    case K_BOOL: v.i = op == R_MAX ? 0 : 1; break; /* bool folds through the int64 slot */
  }
  return v;
}
static void red_combine(int op, int dt, val *m, val v) { /* nx_c_fold.c:63-91 */
  switch (DT_KIND[dt]) {
#define FL(M, V)                                                         \
  if (op == R_SUM) M += V; else if (op == R_PROD) M *= V;                \
  else if (op == R_MAX) { if (V > M) M = V; else if (V != V) M = V; }    \
  else { if (V < M) M = V; else if (V != V) M = V; }
    case K_F32: FL(m->f, v.f) break;
    case K_F64: FL(m->d, v.d) break;
#undef FL
    case K_I64:
      if (op == R_SUM) m->i = (int64_t)((uint64_t)m->i + (uint64_t)v.i);
      else if (op == R_PROD) m->i = (int64_t)((uint64_t)m->i * (uint64_t)v.i);
      else if (op == R_MAX) { if (v.i > m->i) m->i = v.i; }
      else { if (v.i < m->i) m->i = v.i; }
      break;
    case K_U64:
      if (op == R_SUM) m->u += v.u; else if (op == R_PROD) m->u *= v.u;
      else if (op == R_MAX) { if (v.u > m->u) m->u = v.u; }
      else { if (v.u < m->u) m->u = v.u; }
      break;
    case K_C32: if (op == R_SUM) m->c32 += v.c32; else m->c32 *= v.c32; break;
    case K_C64: if (op == R_SUM) m->c64 += v.c64; else m->c64 *= v.c64; break;
    case K_BOOL: if (op == R_MAX) m->i |= v.b; else m->i &= v.b; break;
  }
}
/* one strided run folded into acc; float sums use 16 partials on a contiguous run
   and a fixed balanced tree (nx_c_fold.c:139-159) */
static void red_run(int op, int dt, val *acc, const char *in, int64_t step, int64_t n) {
  int k = DT_KIND[dt];
  if (op == R_SUM && (k == K_F32 || k == K_F64)) {
    if (k == K_F32) {
      float s[16] = {0};
      int64_t i = 0;
      if (step == DT_SIZE[dt]) {
        for (; i + 16 <= n; i += 16) for (int j = 0; j < 16; j++) s[j] += ld(dt, in + (i + j) * step).f;
      }
      for (; i < n; i++) s[0] += ld(dt, in + i * step).f;
      float lo = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
      float hi = ((s[8] + s[9]) + (s[10] + s[11])) + ((s[12] + s[13]) + (s[14] + s[15]));
      acc->f += lo + hi;
    } else {
      double s[16] = {0};
      int64_t i = 0;
      if (step == DT_SIZE[dt]) {
        for (; i + 16 <= n; i += 16) for (int j = 0; j < 16; j++) s[j] += ld(dt, in + (i + j) * step).d;
      }
      for (; i < n; i++) s[0] += ld(dt, in + i * step).d;
      double lo = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
      double hi = ((s[8] + s[9]) + (s[10] + s[11])) + ((s[12] + s[13]) + (s[14] + s[15]));
      acc->d += lo + hi;
    }
    return;
  }
  for (int64_t i = 0; i < n; i++) red_combine(op, dt, acc, ld(dt, in + i * step));
}
static void red_store(int dt, char *p, val acc) {
  if (DT_KIND[dt] == K_BOOL) { val b; b.b = (uint8_t)(acc.i != 0); st(dt, p, b); } else st(dt, p, acc);
}

/* squeeze the out descriptor (nx_c_engine.c:1392-1420) */
static nxo_status squeeze(const nxo_tensor *in, const nxo_tensor *out, const int *axes, int n, int64_t *ostride,
                          int *red) {
  if (n < 0 || n > in->ndim) return E_AXES;
  for (int a = 0; a < in->ndim; a++) red[a] = 0;
  for (int i = 0; i < n; i++) {
    int a = axes[i];
    if (a < 0 || a >= in->ndim || red[a]) return E_AXES;
    red[a] = 1;
  }
  int kept = in->ndim - n;
  if (out->ndim == kept) {
    for (int j = 0; j < kept; j++) ostride[j] = out->strides[j];
  } else if (out->ndim == in->ndim) {
    int j = 0;
    for (int a = 0; a < in->ndim; a++) if (!red[a]) ostride[j++] = out->strides[a];
  } else return E_OUT_RANK;
  return NULL;
}

nxo_status nxo_reduce(int op, const nxo_tensor *out, const nxo_tensor *in, const int *axes, int n_axes) {
  nxo_status s;
  if ((s = chk(in)) || (s = chk(out))) return s;
  int dt = in->dtype, cat = DT_CAT[dt];
  if (op < 0 || op > R_MIN) return E_BAD_OP;
  int64_t ostride[NXO_MAX_NDIM];
  int red[NXO_MAX_NDIM];
  if ((s = squeeze(in, out, axes, n_axes, ostride, red))) return s;
  if (cat & CAT_PACKED) return E_PACKED;
  if (((op == R_SUM || op == R_PROD) && (cat & CAT_BOOL)) || ((op == R_MAX || op == R_MIN) && (cat & CAT_COMPLEX)))
    return E_UNSUPPORTED;
  for (int i = 1; i < n_axes; i++) if (axes[i] <= axes[i - 1]) return E_AXES;
  int64_t es = DT_SIZE[dt];
  /* kept and reduced dims; the smallest-|stride| reduced axis becomes the run
     (nx_c_engine.c:1151-1164) */
  int nk = 0, nr = 0;
  int64_t ks[NXO_MAX_NDIM], kin[NXO_MAX_NDIM], kout[NXO_MAX_NDIM], rs[NXO_MAX_NDIM], rin[NXO_MAX_NDIM];
  int64_t O = 1, R = 1;
  for (int a = 0, j = 0; a < in->ndim; a++) {
    if (red[a]) { rs[nr] = in->shape[a]; rin[nr] = in->strides[a] * es; R *= in->shape[a]; nr++; }
    else { ks[nk] = in->shape[a]; kin[nk] = in->strides[a] * es; kout[nk] = ostride[j++] * es; O *= in->shape[a]; nk++; }
  }
  if (O == 0) return NULL;
  if ((op == R_MAX || op == R_MIN) && R == 0) return E_EMPTY_REDUCE;
  if (nr > 1) {
    int best = 0;
    for (int r = 1; r < nr; r++) if (llabs(rin[r]) < llabs(rin[best])) best = r;
    int64_t t = rs[best]; rs[best] = rs[nr - 1]; rs[nr - 1] = t;
    t = rin[best]; rin[best] = rin[nr - 1]; rin[nr - 1] = t;
  }
  const char *ibase = (const char *)in->data + in->offset * es;
  char *obase = (char *)out->data + out->offset * es;
  int64_t kc[NXO_MAX_NDIM] = {0};
  const char *ip = ibase;
  char *opp = obase;
  for (int64_t o = 0; o < O; o++) {
    val acc = red_init(op, dt);
    if (nr == 0) red_run(op, dt, &acc, ip, 0, 1);
    else {
      int64_t outer = 1, rc[NXO_MAX_NDIM] = {0};
      for (int d = 0; d < nr - 1; d++) outer *= rs[d];
      const char *rp = ip;
      for (int64_t q = 0; q < outer; q++) {
        red_run(op, dt, &acc, rp, rin[nr - 1], rs[nr - 1]);
        for (int d = nr - 2; d >= 0; d--) {
          if (++rc[d] < rs[d]) { rp += rin[d]; break; }
          rc[d] = 0;
          rp -= (rs[d] - 1) * rin[d];
        }
      }
    }
    red_store(dt, opp, acc);
    for (int d = nk - 1; d >= 0; d--) {
      if (++kc[d] < ks[d]) { ip += kin[d]; opp += kout[d]; break; }
      kc[d] = 0;
      ip -= (ks[d] - 1) * kin[d];
      opp -= (ks[d] - 1) * kout[d];
    }
  }
  return NULL;
}

/* argmax/argmin: strict compare keeps the first index, the first NaN wins
   (nx_c_fold.c:93-101, 180-198) */
static int arg_better(int is_max, int dt, val v, val b) {
  switch (DT_KIND[dt]) {
    case K_F32: return is_max ? ((v.f > b.f) || ((v.f != v.f) && !(b.f != b.f))) : ((v.f < b.f) || ((v.f != v.f) && !(b.f != b.f)));
    case K_F64: return is_max ? ((v.d > b.d) || ((v.d != v.d) && !(b.d != b.d))) : ((v.d < b.d) || ((v.d != v.d) && !(b.d != b.d)));
    case K_I64: return is_max ? v.i > b.i : v.i < b.i;
    case K_U64: return is_max ? v.u > b.u : v.u < b.u;
    case K_BOOL: return is_max ? v.b > b.b : v.b < b.b;
  }
  return 0;
}
nxo_status nxo_argreduce(int is_max, const nxo_tensor *out, const nxo_tensor *in, int axis) {
  nxo_status s;
  if ((s = chk(in)) || (s = chk(out))) return s;
  int dt = in->dtype;
  int64_t ostride[NXO_MAX_NDIM];
  int red[NXO_MAX_NDIM];
  if ((s = squeeze(in, out, &axis, 1, ostride, red))) return s;
  if (DT_CAT[dt] & CAT_PACKED) return E_PACKED;
  if (DT_CAT[dt] & CAT_COMPLEX) return E_UNSUPPORTED;
  int64_t len = in->shape[axis];
  if (len == 0) return E_EMPTY_REDUCE;
  if (len > INT32_MAX) return E_ARGCAP;
  int64_t es = DT_SIZE[dt];
  int nk = 0;
  int64_t ks[NXO_MAX_NDIM], kin[NXO_MAX_NDIM], kout[NXO_MAX_NDIM], O = 1;
  for (int a = 0; a < in->ndim; a++) {
    if (a == axis) continue;
    ks[nk] = in->shape[a]; kin[nk] = in->strides[a] * es; kout[nk] = ostride[nk] * 4; O *= in->shape[a]; nk++;
  }
  const char *ip = (const char *)in->data + in->offset * es;
  char *opp = (char *)out->data + out->offset * 4;
  int64_t kc[NXO_MAX_NDIM] = {0}, astep = in->strides[axis] * es;
  for (int64_t o = 0; o < O; o++) {
    val best = ld(dt, ip);
    int64_t bi = 0;
    for (int64_t k = 1; k < len; k++) {
      val v = ld(dt, ip + k * astep);
      if (arg_better(is_max, dt, v, best)) { best = v; bi = k; }
    }
    *(int32_t *)opp = (int32_t)bi;
    for (int d = nk - 1; d >= 0; d--) {
      if (++kc[d] < ks[d]) { ip += kin[d]; opp += kout[d]; break; }
      kc[d] = 0;
      ip -= (ks[d] - 1) * kin[d];
      opp -= (ks[d] - 1) * kout[d];
    }
  }
  return NULL;
}

/* inclusive scan along one axis (nx_c_fold.c:161-176, nx_c_engine.c:1293-1331) */
nxo_status nxo_scan(int op, const nxo_tensor *out, const nxo_tensor *in, int axis) {
  nxo_status s;
  if ((s = chk(in)) || (s = chk(out))) return s;
  int dt = in->dtype, cat = DT_CAT[dt];
  if (cat & CAT_PACKED) return E_PACKED;
  if (((op == R_SUM || op == R_PROD) && (cat & CAT_BOOL)) || ((op == R_MAX || op == R_MIN) && (cat & CAT_COMPLEX)))
    return E_UNSUPPORTED;
  if (axis < 0 || axis >= in->ndim) return E_AXIS;
  if (out->ndim != in->ndim) return E_OUT_RANK;
  int64_t es = DT_SIZE[dt], len = in->shape[axis];
  int nk = 0;
  int64_t ks[NXO_MAX_NDIM], kin[NXO_MAX_NDIM], kout[NXO_MAX_NDIM], O = 1;
  for (int a = 0; a < in->ndim; a++) {
    if (a == axis) continue;
    ks[nk] = in->shape[a]; kin[nk] = in->strides[a] * es; kout[nk] = out->strides[a] * es; O *= in->shape[a]; nk++;
  }
  if (O == 0 || len == 0) return NULL;
  const char *ip = (const char *)in->data + in->offset * es;
  char *opp = (char *)out->data + out->offset * es;
  int64_t kc[NXO_MAX_NDIM] = {0}, ai = in->strides[axis] * es, ao = out->strides[axis] * es;
  for (int64_t o = 0; o < O; o++) {
    val acc = red_init(op, dt);
    for (int64_t k = 0; k < len; k++) {
      red_combine(op, dt, &acc, ld(dt, ip + k * ai));
      red_store(dt, opp + k * ao, acc);
    }
    for (int d = nk - 1; d >= 0; d--) {
      if (++kc[d] < ks[d]) { ip += kin[d]; opp += kout[d]; break; }
      kc[d] = 0;
      ip -= (ks[d] - 1) * kin[d];
      opp -= (ks[d] - 1) * kout[d];
    }
  }
  return NULL;
}

/* ==== matmul (nx_c_matmul.c:874-936; accumulate in the compute type, store once) ======= */
nxo_status nxo_matmul(const nxo_tensor *C, const nxo_tensor *A, const nxo_tensor *B) {
  nxo_status s;
  if ((s = chk(A)) || (s = chk(B)) || (s = chk(C))) return s;
  if (A->dtype != B->dtype || A->dtype != C->dtype) return E_MM_DTYPE;
  int dt = A->dtype;
  if (DT_CAT[dt] & CAT_PACKED) return E_PACKED;
  if (DT_CAT[dt] & CAT_BOOL) return E_UNSUPPORTED;
  if (A->ndim < 2 || B->ndim < 2) return E_SHAPE;
  int nd = A->ndim > B->ndim ? A->ndim : B->ndim;
  if (C->ndim != nd) return E_SHAPE;
  int64_t m = A->shape[A->ndim - 2], k = A->shape[A->ndim - 1], n = B->shape[B->ndim - 1];
  if (k != B->shape[B->ndim - 2]) return E_SHAPE;
  if (C->shape[nd - 2] != m || C->shape[nd - 1] != n) return E_SHAPE;
  int bnd = nd - 2, a_bo = nd - A->ndim, b_bo = nd - B->ndim;
  int64_t bshape[NXO_MAX_NDIM], as_[NXO_MAX_NDIM], bs_[NXO_MAX_NDIM], cs_[NXO_MAX_NDIM], nb = 1;
  for (int i = 0; i < bnd; i++) {
    int64_t sa = 1, sb = 1, sta = 0, stb = 0;
    if (i >= a_bo) { sa = A->shape[i - a_bo]; sta = A->strides[i - a_bo]; }
    if (i >= b_bo) { sb = B->shape[i - b_bo]; stb = B->strides[i - b_bo]; }
    if (sa != sb && sa != 1 && sb != 1) return E_SHAPE;
    int64_t sz = sa > sb ? sa : sb;
    if (C->shape[i] != sz) return E_SHAPE;
    bshape[i] = sz; as_[i] = sa == 1 ? 0 : sta; bs_[i] = sb == 1 ? 0 : stb; cs_[i] = C->strides[i];
    if (sz > 1 && cs_[i] == 0) return E_ALIASED;
    nb *= sz;
  }
  if (m == 0 || n == 0 || nb == 0) return NULL;
  int64_t es = DT_SIZE[dt];
  int64_t a_rs = A->strides[A->ndim - 2], a_cs = A->strides[A->ndim - 1];
  int64_t b_rs = B->strides[B->ndim - 2], b_cs = B->strides[B->ndim - 1];
  int64_t c_rs = C->strides[nd - 2], c_cs = C->strides[nd - 1];
  if ((m > 1 && c_rs == 0) || (n > 1 && c_cs == 0)) return E_ALIASED;
  int64_t bc[NXO_MAX_NDIM] = {0};
  for (int64_t bt = 0; bt < nb; bt++) {
    int64_t ao = A->offset, bo = B->offset, co = C->offset;
    for (int i = 0; i < bnd; i++) { ao += bc[i] * as_[i]; bo += bc[i] * bs_[i]; co += bc[i] * cs_[i]; }
    const char *ap = (const char *)A->data + ao * es, *bp = (const char *)B->data + bo * es;
    char *cp = (char *)C->data + co * es;
    for (int64_t i = 0; i < m; i++)
      for (int64_t j = 0; j < n; j++) {
        val acc = red_init(R_SUM, dt);
        for (int64_t p = 0; p < k; p++) {
          val x = ld(dt, ap + (i * a_rs + p * a_cs) * es), y = ld(dt, bp + (p * b_rs + j * b_cs) * es);
          red_combine(R_SUM, dt, &acc, bin_apply(MUL, dt, x, y));
        }
        st(dt, cp + (i * c_rs + j * c_cs) * es, acc);
      }
    for (int d = bnd - 1; d >= 0; d--) { if (++bc[d] < bshape[d]) break; bc[d] = 0; }
  }
  return NULL;
}

/* ==== pad / cat / gather / scatter (nx_c_move.c:229-569) ================================ */
nxo_status nxo_fill(const nxo_tensor *out, const void *scalar) {
  nxo_status s;
  if ((s = chk(out))) return s;
  int es = DT_SIZE[out->dtype];
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) memcpy(o.ptr[0], scalar, (size_t)es);
  return NULL;
}
nxo_status nxo_pad(const nxo_tensor *out, const nxo_tensor *in, const void *fill, const int64_t *before) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(in))) return s;
  if (DT_CAT[out->dtype] & CAT_PACKED) return E_PACKED;
  if (out->ndim != in->ndim) return E_SHAPE;
  if ((s = nxo_fill(out, fill))) return s;
  nxo_tensor win = *out;
  for (int i = 0; i < in->ndim; i++) {
    if (before[i] < 0 || before[i] + in->shape[i] > out->shape[i]) return E_SHAPE;
    win.offset += before[i] * out->strides[i];
    win.shape[i] = in->shape[i];
  }
  return nxo_copy(&win, in);
}
nxo_status nxo_cat(const nxo_tensor *out, const nxo_tensor *const *ins, int n, int axis) {
  nxo_status s;
  if ((s = chk(out))) return s;
  if (axis < 0 || axis >= out->ndim) return E_AXIS;
  int64_t at = 0;
  for (int t = 0; t < n; t++) {
    if ((s = chk(ins[t]))) return s;
    if (ins[t]->ndim != out->ndim) return E_SHAPE;
    nxo_tensor win = *out;
    win.offset += at * out->strides[axis];
    for (int i = 0; i < out->ndim; i++) {
      if (i != axis && ins[t]->shape[i] != out->shape[i]) return E_SHAPE;
      win.shape[i] = ins[t]->shape[i];
    }
    if ((s = nxo_copy(&win, ins[t]))) return s;
    at += ins[t]->shape[axis];
  }
  if (at != out->shape[axis]) return E_SHAPE;
  return NULL;
}
/* out[c] = data[c with axis -> idx[c]]; indices are int32, Python-wrapped once, then
   bounds-checked: an out-of-range index is an error, never clamped
   (nx_c_move.c:342-362, 402-442) */
static int norm_index(int64_t *ix, int64_t len) {
  if (*ix < 0) *ix += len;
  return *ix >= 0 && *ix < len;
}
nxo_status nxo_gather(const nxo_tensor *out, const nxo_tensor *data, const nxo_tensor *idx, int axis) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(data)) || (s = chk(idx))) return s;
  if (DT_CAT[out->dtype] & CAT_PACKED) return E_PACKED;
  if (axis < 0 || axis >= data->ndim) return E_AXIS;
  if (idx->ndim != data->ndim || out->ndim != data->ndim) return E_SHAPE;
  for (int d = 0; d < out->ndim; d++) if (out->shape[d] != idx->shape[d]) return E_SHAPE;
  int64_t es = DT_SIZE[data->dtype], len = data->shape[axis];
  odo o;
  odo_init(&o, out->ndim, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, idx, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) {
    int64_t ix = *(int32_t *)o.ptr[1];
    if (!norm_index(&ix, len)) return E_INDEX_OOB;
    int64_t off = data->offset;
    for (int d = 0; d < data->ndim; d++) off += (d == axis ? ix : o.coord[d]) * data->strides[d];
    memcpy(o.ptr[0], (const char *)data->data + off * es, (size_t)es);
  }
  return NULL;
}
/* `Set: the last write in row-major order wins; `Add accumulates in the compute type,
   serially in row-major order (nx_c_move.c:444-569) */
nxo_status nxo_scatter(const nxo_tensor *out, const nxo_tensor *idx, const nxo_tensor *upd, int axis, int mode) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(upd)) || (s = chk(idx))) return s;
  int dt = out->dtype;
  if (DT_CAT[dt] & CAT_PACKED) return E_PACKED;
  if (axis < 0 || axis >= out->ndim) return E_AXIS;
  if (out->ndim != idx->ndim || out->ndim != upd->ndim) return E_SHAPE;
  for (int d = 0; d < out->ndim; d++) {
    if (idx->shape[d] != upd->shape[d]) return E_SHAPE;
    if (d != axis && idx->shape[d] != out->shape[d]) return E_SHAPE;
  }
  int64_t es = DT_SIZE[dt], len = out->shape[axis];
  odo o;
  odo_init(&o, idx->ndim, idx->shape);
  odo_add(&o, idx, NULL);
  odo_add(&o, upd, NULL);
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) {
    int64_t ix = *(int32_t *)o.ptr[0];
    if (!norm_index(&ix, len)) return E_INDEX_OOB;
    int64_t off = out->offset;
    for (int d = 0; d < out->ndim; d++) off += (d == axis ? ix : o.coord[d]) * out->strides[d];
    char *dst = (char *)out->data + off * es;
    if (mode == 0) memcpy(dst, o.ptr[1], (size_t)es);
    else if (DT_KIND[dt] == K_BOOL) { val r; r.b = (uint8_t)((ld(dt, dst).b + ld(dt, o.ptr[1]).b) != 0); st(dt, dst, r); }
    else st(dt, dst, bin_apply(ADD, dt, ld(dt, dst), ld(dt, o.ptr[1])));
  }
  return NULL;
}

/* ==== unfold / fold (nx_c_move.c:588-870) =================================================
   Window geometry: per spatial dim, win = max(1, (extent + pad_before + pad_after - eff) / stride + 1)
   with eff = dilation * (kernel - 1) + 1 (nx_c_move.c:632-646). Restated in the FORWARD direction:
   walk (leading, kernel offset, window) and visit the one source position each pair names.
     unfold  copies that position into out[lead, kf, wf], or zero when it falls in the padding
             (nx_c_move.c:660-696).
     fold    adds in[lead, kf, wf] into a compute-typed accumulator of that position. kf is the
             OUTER loop, so every output still receives its taps in ascending-kf order -- the
             order the reference's per-output gather sums them in (nx_c_move.c:755-790) -- and
             the accumulator is rounded to storage once, at the end. */
typedef struct {
  int K;
  int64_t kernel[NXO_MAX_NDIM], stride[NXO_MAX_NDIM], dil[NXO_MAX_NDIM], before[NXO_MAX_NDIM];
  int64_t extent[NXO_MAX_NDIM], win[NXO_MAX_NDIM];
  int64_t kprod, nwin, nspatial;
} wgeom;
static nxo_status wgeom_init(wgeom *g, int K, const int64_t *extent, const int64_t *kernel, const int64_t *stride,
                             const int64_t *dil, const int64_t *pad) {
  if (K < 1 || K > NXO_MAX_NDIM) return E_SHAPE;
  g->K = K; g->kprod = 1; g->nwin = 1; g->nspatial = 1;
  for (int d = 0; d < K; d++) {
    g->kernel[d] = kernel[d]; g->stride[d] = stride[d]; g->dil[d] = dil[d]; g->before[d] = pad[2 * d];
    g->extent[d] = extent[d];
    int64_t span = dil[d] * (kernel[d] - 1) + 1, w = (extent[d] + pad[2 * d] + pad[2 * d + 1] - span) / stride[d] + 1;
    g->win[d] = w < 1 ? 1 : w;
    g->kprod *= kernel[d]; g->nwin *= g->win[d]; g->nspatial *= extent[d];
  }
  return NULL;
}
/* position (row-major over extent) named by kernel offset kc and window wc, or -1 in the padding;
   *off gets the strided element offset over `strides` */
static int64_t wgeom_pos(const wgeom *g, const int64_t *kc, const int64_t *wc, const int64_t *strides, int64_t *off) {
  int64_t lin = 0;
  *off = 0;
  for (int d = 0; d < g->K; d++) {
    int64_t sp = wc[d] * g->stride[d] + kc[d] * g->dil[d] - g->before[d];
    if (sp < 0 || sp >= g->extent[d]) return -1;
    lin = lin * g->extent[d] + sp;
    *off += sp * strides[d];
  }
  return lin;
}
static void unravel_rm(int64_t lin, int n, const int64_t *shape, int64_t *coord) {
  for (int d = n - 1; d >= 0; d--) { coord[d] = shape[d] ? lin % shape[d] : 0; if (shape[d]) lin /= shape[d]; }
}
nxo_status nxo_unfold(const nxo_tensor *out, const nxo_tensor *in, int K, const int64_t *kernel,
                      const int64_t *stride, const int64_t *dil, const int64_t *pad) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(in))) return s;
  if (DT_CAT[out->dtype] & CAT_PACKED) return E_PACKED;
  if (K < 1 || in->ndim < K) return E_SHAPE;
  int ld = in->ndim - K;
  wgeom g;
  if ((s = wgeom_init(&g, K, &in->shape[ld], kernel, stride, dil, pad))) return s;
  int64_t es = DT_SIZE[out->dtype], nlead = 1, kp = out->shape[ld], L = out->shape[ld + 1];
  for (int d = 0; d < ld; d++) nlead *= in->shape[d];
  int64_t lc[NXO_MAX_NDIM], kc[NXO_MAX_NDIM], wc[NXO_MAX_NDIM];
  for (int64_t l = 0; l < nlead; l++) {
    unravel_rm(l, ld, in->shape, lc);
    int64_t ibase = in->offset, obase = out->offset;
    for (int d = 0; d < ld; d++) { ibase += lc[d] * in->strides[d]; obase += lc[d] * out->strides[d]; }
    for (int64_t kf = 0; kf < kp; kf++) {
      unravel_rm(kf, K, g.kernel, kc);
      for (int64_t wf = 0; wf < L; wf++) {
        unravel_rm(wf, K, g.win, wc);
        int64_t off;
        char *dst = (char *)out->data + (obase + kf * out->strides[ld] + wf * out->strides[ld + 1]) * es;
        if (wgeom_pos(&g, kc, wc, &in->strides[ld], &off) < 0) memset(dst, 0, (size_t)es);
        else memcpy(dst, (const char *)in->data + (ibase + off) * es, (size_t)es);
      }
    }
  }
  return NULL;
}
nxo_status nxo_fold(const nxo_tensor *out, const nxo_tensor *in, int K, const int64_t *output_size,
                    const int64_t *kernel, const int64_t *stride, const int64_t *dil, const int64_t *pad) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(in))) return s;
  int dt = out->dtype;
  if (DT_CAT[dt] & CAT_PACKED) return E_PACKED;
  if (K < 1 || in->ndim < 2) return E_SHAPE;
  int nl = in->ndim - 2;
  wgeom g;
  if ((s = wgeom_init(&g, K, output_size, kernel, stride, dil, pad))) return s;
  int64_t es = DT_SIZE[dt], nlead = 1, kp = in->shape[nl];
  for (int d = 0; d < nl; d++) nlead *= in->shape[d];
  if (nlead * g.nspatial == 0) return NULL;
  /* the reference trusts the frontend here and would read past the column tensor; the
     restatement (and the device engine) refuse instead */
  if (in->shape[nl + 1] != g.nwin || kp != g.kprod) return E_SHAPE;
  val *acc = (val *)malloc(sizeof(val) * (size_t)g.nspatial);
  int64_t lc[NXO_MAX_NDIM], kc[NXO_MAX_NDIM], wc[NXO_MAX_NDIM], oc[NXO_MAX_NDIM];
  for (int64_t l = 0; l < nlead; l++) {
    unravel_rm(l, nl, in->shape, lc);
    int64_t ibase = in->offset, obase = out->offset;
    for (int d = 0; d < nl; d++) { ibase += lc[d] * in->strides[d]; obase += lc[d] * out->strides[d]; }
    memset(acc, 0, sizeof(val) * (size_t)g.nspatial);
    for (int64_t kf = 0; kf < kp; kf++) {
      unravel_rm(kf, K, g.kernel, kc);
      for (int64_t wf = 0; wf < g.nwin; wf++) {
        unravel_rm(wf, K, g.win, wc);
        int64_t off, pos = wgeom_pos(&g, kc, wc, &out->strides[nl], &off);
        if (pos < 0) continue;
        val v = ld(dt, (const char *)in->data + (ibase + kf * in->strides[nl] + wf * in->strides[nl + 1]) * es);
        if (DT_KIND[dt] == K_BOOL) acc[pos].b = (uint8_t)(acc[pos].b + v.b); /* uint8_t compute type */
        else acc[pos] = bin_apply(ADD, dt, acc[pos], v);
      }
    }
    for (int64_t p = 0; p < g.nspatial; p++) {
      unravel_rm(p, K, g.extent, oc);
      int64_t off = obase;
      for (int d = 0; d < K; d++) off += oc[d] * out->strides[nl + d];
      st(dt, (char *)out->data + off * es, acc[p]);
    }
  }
  free(acc);
  return NULL;
}

/* ==== threefry2x32, 20 rounds (nx_c_random.c:44-61; Random123 reference constants) ====== */
static uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t *o0, uint32_t *o1) {
  static const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, 0x1BD11BDAu ^ k0 ^ k1};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int r = 0; r < 20; r++) {
    x0 += x1;
    x1 = rotl(x1, R[r % 8]);
    x1 ^= x0;
    if ((r & 3) == 3) {
      int q = r / 4 + 1;
      x0 += ks[q % 3];
      x1 += ks[(q + 1) % 3] + (uint32_t)q;
    }
  }
  *o0 = x0;
  *o1 = x1;
}
/* key, ctr, out: int32 with a trailing dim of 2 (pairs) */
nxo_status nxo_threefry(const nxo_tensor *out, const nxo_tensor *key, const nxo_tensor *ctr) {
  nxo_status s;
  if ((s = chk(out)) || (s = chk(key)) || (s = chk(ctr))) return s;
  if (out->dtype != I32 || key->dtype != I32 || ctr->dtype != I32) return E_UNSUPPORTED;
  int nd = key->ndim;
  if (nd < 1 || nd != ctr->ndim || nd != out->ndim) return E_THREEFRY_SHAPE;
  for (int d = 0; d < nd; d++)
    if (key->shape[d] != ctr->shape[d] || key->shape[d] != out->shape[d]) return E_THREEFRY_SHAPE;
  if (key->shape[nd - 1] != 2) return E_THREEFRY_SHAPE;
  odo o;
  odo_init(&o, nd - 1, out->shape);
  odo_add(&o, out, NULL);
  odo_add(&o, key, NULL);
  odo_add(&o, ctr, NULL);
  int64_t so = out->strides[nd - 1] * 4, sk = key->strides[nd - 1] * 4, sc = ctr->strides[nd - 1] * 4;
  for (int64_t i = 0; i < o.total; i++, odo_next(&o)) {
    uint32_t r0, r1;
    threefry2x32(*(uint32_t *)o.ptr[1], *(uint32_t *)(o.ptr[1] + sk), *(uint32_t *)o.ptr[2],
                 *(uint32_t *)(o.ptr[2] + sc), &r0, &r1);
    *(uint32_t *)o.ptr[0] = r0;
    *(uint32_t *)(o.ptr[0] + so) = r1;
  }
  return NULL;
}


/* ==== sort / argsort (nx_c_sort.c:47-101, 265-276) ======================================
   NaN-class elements (any NaN part for complex) last in both directions, in original order;
   complex lexicographic; argsort stable (value ties keep the first index in either
   direction). Restated as a merge sort on (value, original index) under that total order. */
typedef struct { val v; int32_t i; } sort_item;
static int sort_dt, sort_desc;
static int sort_isnan(val x) {
  switch (DT_KIND[sort_dt]) {
    case K_F32: return isnan(x.f);
    case K_F64: return isnan(x.d);
    case K_C32: return isnan(crealf(x.c32)) || isnan(cimagf(x.c32));
    case K_C64: return isnan(creal(x.c64)) || isnan(cimag(x.c64));
  }
  return 0;
}
static int sort_lt(val x, val y) {
  switch (DT_KIND[sort_dt]) {
    case K_F32: return x.f < y.f;
    case K_F64: return x.d < y.d;
    case K_I64: return x.i < y.i;
    case K_U64: return x.u < y.u;
    case K_BOOL: return x.b < y.b;
    case K_C32: return crealf(x.c32) < crealf(y.c32) || (crealf(x.c32) == crealf(y.c32) && cimagf(x.c32) < cimagf(y.c32));
    case K_C64: return creal(x.c64) < creal(y.c64) || (creal(x.c64) == creal(y.c64) && cimag(x.c64) < cimag(y.c64));
  }
  return 0;
}
static int sort_before(const sort_item *a, const sort_item *b) {
  int na = sort_isnan(a->v), nb = sort_isnan(b->v);
  if (na || nb) return (na && nb) ? a->i < b->i : nb;
  if (!sort_lt(a->v, b->v) && !sort_lt(b->v, a->v)) return a->i < b->i;
  return sort_desc ? sort_lt(b->v, a->v) : sort_lt(a->v, b->v);
}
static void sort_merge(sort_item *a, sort_item *tmp, int64_t n) {
  if (n < 2) return;
  int64_t h = n / 2;
  sort_merge(a, tmp, h);
  sort_merge(a + h, tmp, n - h);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = sort_before(&a[j], &a[i]) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof *a);
}
nxo_status nxo_sort(int is_arg, const nxo_tensor *out, const nxo_tensor *in, int axis, int desc) {
  nxo_status s;
  if ((s = chk(in)) || (s = chk(out))) return s;
  int dt = in->dtype;
  if (DT_CAT[dt] & CAT_PACKED) return E_PACKED;
  if (axis < 0 || axis >= in->ndim) return E_AXIS;
  if (out->ndim != in->ndim) return E_OUT_RANK;
  int64_t es = DT_SIZE[dt], oes = is_arg ? 4 : es, len = in->shape[axis];
  int nk = 0;
  int64_t ks[NXO_MAX_NDIM], kin[NXO_MAX_NDIM], kout[NXO_MAX_NDIM], O = 1;
  for (int a = 0; a < in->ndim; a++) {
    if (a == axis) continue;
    ks[nk] = in->shape[a]; kin[nk] = in->strides[a] * es; kout[nk] = out->strides[a] * oes; O *= in->shape[a]; nk++;
  }
  if (O == 0 || len == 0) return NULL;
  for (int a = 0; a < in->ndim; a++) if (in->shape[a] > 1 && out->strides[a] == 0) return E_ALIASED;
  sort_item *buf = malloc((size_t)len * sizeof *buf * 2);
  if (!buf) return "out of memory";
  sort_dt = dt; sort_desc = desc;
  const char *ip = (const char *)in->data + in->offset * es;
  char *opp = (char *)out->data + out->offset * oes;
  int64_t kc[NXO_MAX_NDIM] = {0}, ai = in->strides[axis] * es, ao = out->strides[axis] * oes;
  for (int64_t o = 0; o < O; o++) {
    for (int64_t k = 0; k < len; k++) { buf[k].v = ld(dt, ip + k * ai); buf[k].i = (int32_t)k; }
    sort_merge(buf, buf + len, len);
    for (int64_t k = 0; k < len; k++) {
      if (is_arg) *(int32_t *)(opp + k * ao) = buf[k].i;
      else memcpy(opp + k * ao, ip + buf[k].i * ai, (size_t)es);
    }
    for (int d = nk - 1; d >= 0; d--) {
      if (++kc[d] < ks[d]) { ip += kin[d]; opp += kout[d]; break; }
      kc[d] = 0;
      ip -= (ks[d] - 1) * kin[d];
      opp -= (ks[d] - 1) * kout[d];
    }
  }
  free(buf);
  return NULL;
}

/* 1 -> Invalid_argument, 0 -> Failure (nx_c_engine.c:1345-1351, nx_c_matmul.c:1229-1237) */
int nxo_status_is_invalid_argument(nxo_status s) {
  if (!s) return 0;
  return !strcmp(s, E_EMPTY_REDUCE) || !strcmp(s, E_AXES) || !strcmp(s, E_AXIS) || !strcmp(s, E_OUT_RANK) ||
         !strcmp(s, E_ALIASED) || !strcmp(s, E_SHAPE) || !strcmp(s, E_THREEFRY_SHAPE);
}
