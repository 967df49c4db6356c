/* nx_cuda_stubs.c -- OCaml externals over libnxcuda (include/nxcuda.h).
 *
 * One stub per external in nx_backend.ml. Each extracts slots 0-4 of the tensor
 * record into an nxc_tensor (the same slots, in the same order, that the reference
 * reads: packages/nx/lib/backend_c/nx_c.h:47-61, 420-434), calls one nxc_* entry
 * point (an asynchronous launch on the context stream) and maps a non-NULL status to
 * Invalid_argument / Failure "<op>: <status>" exactly as the reference's funnel does
 * (nx_c_engine.c:42-52, 1345-1351). There is no OCaml toolchain in the build image; this file is
 * nevertheless EXECUTED by the GPU suite: tests/test_gpu_ocaml_stubs.py compiles it against the
 * OCaml C API declarations plus a small runtime shim (custom-block allocation, the two raisers) and
 * drives every stub with fabricated OCaml values, as oracle/ref.py does for the reference's stubs.
 */
#include <stdio.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "nxcuda.h"

typedef struct { void *ptr; size_t bytes; nxc_ctx *ctx; } devbuf;
#define Devbuf_val(v) ((devbuf *)Data_custom_val(v))
#define Ctx_val(v) (*(nxc_ctx **)Data_custom_val(v))

static void devbuf_finalize(value v) {
  devbuf *b = Devbuf_val(v);
  if (b->ptr) nxc_free(b->ctx, b->ptr); /* cudaFreeAsync: ordered after in-flight kernels */
}
static struct custom_operations devbuf_ops = {"nx_cuda.devbuf", devbuf_finalize, custom_compare_default,
    custom_hash_default, custom_serialize_default, custom_deserialize_default, custom_compare_ext_default,
    custom_fixed_length_default};
static struct custom_operations ctx_ops = {"nx_cuda.ctx", custom_finalize_default, custom_compare_default,
    custom_hash_default, custom_serialize_default, custom_deserialize_default, custom_compare_ext_default,
    custom_fixed_length_default};

static void raise_status(const char *op, nxc_ctx *ctx, nxc_status s) {
  char buf[640];
  if (strcmp(s, "CUDA error") == 0 || strcmp(s, "NCCL error") == 0)
    snprintf(buf, sizeof buf, "%s: %s [%s]", op, s, nxc_last_error(ctx));
  else
    snprintf(buf, sizeof buf, "%s: %s", op, s);
  if (nxc_status_is_invalid_argument(s)) caml_invalid_argument(buf);
  caml_failwith(buf);
}

/* record slots: 0 buffer, 1 shape, 2 strides, 3 offset, 4 tag, 5 context */
static void tensor_of_value(value v, nxc_tensor *t) {
  value sh = Field(v, 1), st = Field(v, 2);
  int nd = (int)Wosize_val(sh);
  if (nd > NXC_MAX_NDIM) caml_failwith("ndim exceeds NX_C_MAX_NDIM");
  t->data = Devbuf_val(Field(v, 0))->ptr;
  t->dtype = Int_val(Field(v, 4));
  t->ndim = nd;
  for (int i = 0; i < nd; i++) { t->shape[i] = Long_val(Field(sh, i)); t->strides[i] = Long_val(Field(st, i)); }
  t->offset = Long_val(Field(v, 3));
}
#define CTX_OF(v) Ctx_val(Field((v), 5))

CAMLprim value nx_cuda_ctx_create(value unit) {
  CAMLparam1(unit);
  CAMLlocal1(v);
  nxc_ctx *ctx = NULL;
  nxc_status s = nxc_ctx_create(&ctx);
  if (s) raise_status("create_context", NULL, s);
  v = caml_alloc_custom(&ctx_ops, sizeof(nxc_ctx *), 0, 1);
  Ctx_val(v) = ctx;
  CAMLreturn(v);
}
CAMLprim value nx_cuda_alloc(value vctx, value vbytes) {
  CAMLparam2(vctx, vbytes);
  CAMLlocal1(v);
  size_t n = (size_t)Long_val(vbytes);
  v = caml_alloc_custom_mem(&devbuf_ops, sizeof(devbuf), n); /* the GC sees device pressure */
  devbuf *b = Devbuf_val(v);
  b->ptr = NULL; b->bytes = n; b->ctx = Ctx_val(vctx);
  nxc_status s = nxc_alloc(b->ctx, n, &b->ptr);
  if (s) raise_status("buffer", b->ctx, s);
  CAMLreturn(v);
}
CAMLprim value nx_cuda_of_host(value vctx, value vbuf) {
  CAMLparam2(vctx, vbuf);
  CAMLlocal1(v);
  size_t n = caml_ba_byte_size(Caml_ba_array_val(vbuf));
  v = nx_cuda_alloc(vctx, Val_long(n ? n : 16));
  /* nx_cuda_alloc allocates a custom block, so the GC may have run and moved vbuf's block (it is
     rooted: the VALUE was updated). Everything derived from it is fetched only now. The bigarray's
     data lives outside the OCaml heap and does not move across the blocking section. */
  struct caml_ba_array *ba = Caml_ba_array_val(vbuf);
  void *src = ba->data;
  nxc_ctx *ctx = Ctx_val(vctx);
  nxc_status s = nxc_h2d(ctx, Devbuf_val(v)->ptr, src, n);
  if (!s) { caml_enter_blocking_section(); s = nxc_sync(ctx); caml_leave_blocking_section(); }
  if (s) raise_status("from_host", ctx, s);
  CAMLreturn(v);
}
CAMLprim value nx_cuda_to_host(value vctx, value vdev, value vbuf) {
  CAMLparam3(vctx, vdev, vbuf);
  struct caml_ba_array *ba = Caml_ba_array_val(vbuf);
  nxc_ctx *ctx = Ctx_val(vctx);
  void *src = Devbuf_val(vdev)->ptr, *dst = ba->data;
  size_t n = caml_ba_byte_size(ba);
  caml_enter_blocking_section(); /* the only blocking call on the path */
  nxc_status s = nxc_d2h(ctx, dst, src, n);
  caml_leave_blocking_section();
  if (s) raise_status("to_host", ctx, s);
  CAMLreturn(Val_unit);
}

/* op name for the exception text; an out-of-range code (the engine answers "unknown operation
   code") must not index past the table */
#define OP_NAME(table, i) (((i) >= 0 && (size_t)(i) < sizeof(table) / sizeof((table)[0])) ? (table)[i] : "op")
static const char *UN[] = {"neg","recip","abs","sign","sqrt","exp","log","sin","cos","tan","asin","acos","atan",
                           "sinh","cosh","tanh","trunc","ceil","floor","round","erf"};
static const char *BIN[] = {"add","sub","mul","idiv","fdiv","mod","max","min","pow","atan2","xor","or","and","shl","shr"};
static const char *CMP[] = {"cmpeq","cmpne","cmplt","cmple"};
static const char *RED[] = {"reduce_sum","reduce_prod","reduce_max","reduce_min"};
static const char *SCAN[] = {"cumsum","cumprod","cummax","cummin"};

CAMLprim value nx_cuda_map1(value vop, value vout, value va) {
  CAMLparam3(vop, vout, va);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(va, &a);
  nxc_status s = nxc_map1(CTX_OF(vout), Int_val(vop), &o, &a);
  if (s) raise_status(OP_NAME(UN, Int_val(vop)), CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_map2(value vop, value vout, value va, value vb) {
  CAMLparam4(vop, vout, va, vb);
  nxc_tensor o, a, b; tensor_of_value(vout, &o); tensor_of_value(va, &a); tensor_of_value(vb, &b);
  nxc_status s = nxc_map2(CTX_OF(vout), Int_val(vop), &o, &a, &b);
  if (s) raise_status(OP_NAME(BIN, Int_val(vop)), CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_cmp(value vop, value vout, value va, value vb) {
  CAMLparam4(vop, vout, va, vb);
  nxc_tensor o, a, b; tensor_of_value(vout, &o); tensor_of_value(va, &a); tensor_of_value(vb, &b);
  nxc_status s = nxc_cmp(CTX_OF(vout), Int_val(vop), &o, &a, &b);
  if (s) raise_status(OP_NAME(CMP, Int_val(vop)), CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_where(value vout, value vc, value va, value vb) {
  CAMLparam4(vout, vc, va, vb);
  nxc_tensor o, c, a, b;
  tensor_of_value(vout, &o); tensor_of_value(vc, &c); tensor_of_value(va, &a); tensor_of_value(vb, &b);
  nxc_status s = nxc_where(CTX_OF(vout), &o, &c, &a, &b);
  if (s) raise_status("where", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
#define STUB2(name, opname, call)                                             \
  CAMLprim value nx_cuda_##name(value vout, value va) {                       \
    CAMLparam2(vout, va);                                                     \
    nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(va, &a);      \
    nxc_status s = call(CTX_OF(vout), &o, &a);                                \
    if (s) raise_status(opname, CTX_OF(vout), s);                             \
    CAMLreturn(Val_unit);                                                     \
  }
STUB2(cast, "cast", nxc_cast)
STUB2(copy, "copy", nxc_copy)

CAMLprim value nx_cuda_fill(value vout, value vone) {
  CAMLparam2(vout, vone);
  nxc_tensor o; tensor_of_value(vout, &o);
  nxc_status s = nxc_fill(CTX_OF(vout), &o, Caml_ba_array_val(vone)->data);
  if (s) raise_status("full", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_reduce(value vop, value vout, value vin, value vaxes) {
  CAMLparam4(vop, vout, vin, vaxes);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int n = (int)Wosize_val(vaxes), axes[NXC_MAX_NDIM];
  if (n > NXC_MAX_NDIM) caml_failwith("ndim exceeds NX_C_MAX_NDIM");
  for (int i = 0; i < n; i++) axes[i] = Int_val(Field(vaxes, i));
  nxc_status s = nxc_reduce(CTX_OF(vout), Int_val(vop), &o, &a, axes, n);
  if (s) raise_status(OP_NAME(RED, Int_val(vop)), CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_argreduce(value vmax, value vout, value vin, value vaxis) {
  CAMLparam4(vmax, vout, vin, vaxis);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  nxc_status s = nxc_argreduce(CTX_OF(vout), Bool_val(vmax), &o, &a, Int_val(vaxis));
  if (s) raise_status(Bool_val(vmax) ? "argmax" : "argmin", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_scan(value vop, value vout, value vin, value vaxis) {
  CAMLparam4(vop, vout, vin, vaxis);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  nxc_status s = nxc_scan(CTX_OF(vout), Int_val(vop), &o, &a, Int_val(vaxis));
  if (s) raise_status(OP_NAME(SCAN, Int_val(vop)), CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_matmul(value vout, value va, value vb) {
  CAMLparam3(vout, va, vb);
  nxc_tensor o, a, b; tensor_of_value(vout, &o); tensor_of_value(va, &a); tensor_of_value(vb, &b);
  nxc_status s = nxc_matmul(CTX_OF(vout), &o, &a, &b);
  if (s) raise_status("matmul", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_pad(value vout, value vin, value vone, value vbefore) {
  CAMLparam4(vout, vin, vone, vbefore);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int64_t before[NXC_MAX_NDIM];
  if ((int)Wosize_val(vbefore) != o.ndim) caml_invalid_argument("pad: shape mismatch");
  for (int i = 0; i < o.ndim; i++) before[i] = Long_val(Field(vbefore, i));
  nxc_status s = nxc_pad(CTX_OF(vout), &o, &a, Caml_ba_array_val(vone)->data, before);
  if (s) raise_status("pad", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_cat(value vout, value vins, value vaxis) {
  CAMLparam3(vout, vins, vaxis);
  int n = (int)Wosize_val(vins);
  nxc_tensor o; tensor_of_value(vout, &o);
  nxc_tensor *ts = (nxc_tensor *)caml_stat_alloc(sizeof(nxc_tensor) * (n ? n : 1));
  const nxc_tensor **ps = (const nxc_tensor **)caml_stat_alloc(sizeof(void *) * (n ? n : 1));
  for (int i = 0; i < n; i++) { tensor_of_value(Field(vins, i), &ts[i]); ps[i] = &ts[i]; }
  nxc_status s = nxc_cat(CTX_OF(vout), &o, ps, n, Int_val(vaxis));
  caml_stat_free(ts); caml_stat_free(ps);
  if (s) raise_status("cat", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_gather(value vout, value vdata, value vidx, value vaxis) {
  CAMLparam4(vout, vdata, vidx, vaxis);
  nxc_tensor o, d, i; tensor_of_value(vout, &o); tensor_of_value(vdata, &d); tensor_of_value(vidx, &i);
  nxc_status s = nxc_gather(CTX_OF(vout), &o, &d, &i, Int_val(vaxis));
  if (s) raise_status("gather", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_scatter(value vout, value vidx, value vupd, value vaxis, value vmode) {
  CAMLparam5(vout, vidx, vupd, vaxis, vmode);
  nxc_tensor o, i, u; tensor_of_value(vout, &o); tensor_of_value(vidx, &i); tensor_of_value(vupd, &u);
  nxc_status s = nxc_scatter(CTX_OF(vout), &o, &i, &u, Int_val(vaxis), Int_val(vmode));
  if (s) raise_status("scatter", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_threefry(value vout, value vkey, value vctr) {
  CAMLparam3(vout, vkey, vctr);
  nxc_tensor o, k, c; tensor_of_value(vout, &o); tensor_of_value(vkey, &k); tensor_of_value(vctr, &c);
  nxc_status s = nxc_threefry(CTX_OF(vout), &o, &k, &c);
  if (s) raise_status("threefry", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
static void ints_of_value(value v, int64_t *dst, int cap) {
  int n = (int)Wosize_val(v);
  for (int i = 0; i < n && i < cap; i++) dst[i] = Long_val(Field(v, i));
}
CAMLprim value nx_cuda_unfold(value vout, value vin, value vk, value vs, value vd, value vp) {
  CAMLparam5(vout, vin, vk, vs, vd);
  CAMLxparam1(vp);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int64_t k[NXC_MAX_NDIM], st[NXC_MAX_NDIM], d[NXC_MAX_NDIM], p[2 * NXC_MAX_NDIM];
  ints_of_value(vk, k, NXC_MAX_NDIM); ints_of_value(vs, st, NXC_MAX_NDIM);
  ints_of_value(vd, d, NXC_MAX_NDIM); ints_of_value(vp, p, 2 * NXC_MAX_NDIM);
  nxc_status s = nxc_unfold(CTX_OF(vout), &o, &a, (int)Wosize_val(vk), k, st, d, p);
  if (s) raise_status("unfold", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_unfold_bc(value *argv, int argn) {
  (void)argn;
  return nx_cuda_unfold(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5]);
}
CAMLprim value nx_cuda_fold(value vout, value vin, value vsz, value vk, value vs, value vd, value vp) {
  CAMLparam5(vout, vin, vsz, vk, vs);
  CAMLxparam2(vd, vp);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int64_t sz[NXC_MAX_NDIM], k[NXC_MAX_NDIM], st[NXC_MAX_NDIM], d[NXC_MAX_NDIM], p[2 * NXC_MAX_NDIM];
  ints_of_value(vsz, sz, NXC_MAX_NDIM); ints_of_value(vk, k, NXC_MAX_NDIM); ints_of_value(vs, st, NXC_MAX_NDIM);
  ints_of_value(vd, d, NXC_MAX_NDIM); ints_of_value(vp, p, 2 * NXC_MAX_NDIM);
  nxc_status s = nxc_fold(CTX_OF(vout), &o, &a, (int)Wosize_val(vk), sz, k, st, d, p);
  if (s) raise_status("fold", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_fold_bc(value *argv, int argn) {
  (void)argn;
  return nx_cuda_fold(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]);
}
CAMLprim value nx_cuda_sort(value varg, value vout, value vin, value vaxis, value vdesc) {
  CAMLparam5(varg, vout, vin, vaxis, vdesc);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  nxc_status s = nxc_sort(CTX_OF(vout), Bool_val(varg), &o, &a, Int_val(vaxis), Bool_val(vdesc));
  if (s) raise_status(Bool_val(varg) ? "argsort" : "sort", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}

/* ---- fft family (replaces caml_nx_c_fft / _ifft / _rfft / _irfft, nx_c_fft.c:1173-1223) ---- */
static int axes_of_value(value vaxes, int *axes) {
  int n = (int)Wosize_val(vaxes);
  if (n > NXC_MAX_NDIM) n = NXC_MAX_NDIM;
  for (int i = 0; i < n; i++) axes[i] = Int_val(Field(vaxes, i));
  return n;
}
CAMLprim value nx_cuda_fft(value vinv, value vout, value vin, value vaxes) {
  CAMLparam4(vinv, vout, vin, vaxes);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int axes[NXC_MAX_NDIM], n = axes_of_value(vaxes, axes);
  nxc_status s = nxc_fft(CTX_OF(vout), &o, &a, axes, n, Bool_val(vinv));
  if (s) raise_status(Bool_val(vinv) ? "ifft" : "fft", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_rfft(value vout, value vin, value vaxes) {
  CAMLparam3(vout, vin, vaxes);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int axes[NXC_MAX_NDIM], n = axes_of_value(vaxes, axes);
  nxc_status s = nxc_rfft(CTX_OF(vout), &o, &a, axes, n);
  if (s) raise_status("rfft", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_irfft(value vout, value vin, value vaxes, value vs) {
  CAMLparam4(vout, vin, vaxes, vs);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  int axes[NXC_MAX_NDIM], n = axes_of_value(vaxes, axes);
  int ns = (int)Wosize_val(vs);
  int64_t s_last = ns > 0 ? Long_val(Field(vs, ns - 1)) : 0;
  nxc_status s = nxc_irfft(CTX_OF(vout), &o, &a, axes, n, s_last);
  if (s) raise_status("irfft", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}

/* ---- linalg tier 1 (replaces caml_nx_c_cholesky / _triangular_solve, nx_c_tri.c:570-620, and
   caml_nx_c_qr, nx_c_qr.c:421-433) ---- */
CAMLprim value nx_cuda_cholesky(value vout, value vin, value vupper) {
  CAMLparam3(vout, vin, vupper);
  nxc_tensor o, a; tensor_of_value(vout, &o); tensor_of_value(vin, &a);
  nxc_status s = nxc_cholesky(CTX_OF(vout), &o, &a, Bool_val(vupper));
  if (s) raise_status("cholesky", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_triangular_solve(value vout, value va, value vb, value vflags) {
  CAMLparam4(vout, va, vb, vflags);
  nxc_tensor o, a, b; tensor_of_value(vout, &o); tensor_of_value(va, &a); tensor_of_value(vb, &b);
  nxc_status s = nxc_triangular_solve(CTX_OF(vout), &o, &a, &b, Int_val(vflags));
  if (s) raise_status("triangular_solve", CTX_OF(vout), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_qr(value vq, value vr, value vin, value vreduced) {
  CAMLparam4(vq, vr, vin, vreduced);
  nxc_tensor q, r, a; tensor_of_value(vq, &q); tensor_of_value(vr, &r); tensor_of_value(vin, &a);
  nxc_status s = nxc_qr(CTX_OF(vq), &q, &r, &a, Bool_val(vreduced));
  if (s) raise_status("qr", CTX_OF(vq), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_eigh(value vw, value vv, value vin, value vvectors) {
  CAMLparam4(vw, vv, vin, vvectors);
  nxc_tensor w, v, a; tensor_of_value(vw, &w); tensor_of_value(vv, &v); tensor_of_value(vin, &a);
  nxc_status s = nxc_eigh(CTX_OF(vw), &w, &v, &a, Bool_val(vvectors));
  if (s) raise_status(Bool_val(vvectors) ? "eigh" : "eigvalsh", CTX_OF(vw), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_svd(value vu, value vs, value vvt, value vin) {
  CAMLparam4(vu, vs, vvt, vin);
  nxc_tensor u, sv, vt, a; tensor_of_value(vu, &u); tensor_of_value(vs, &sv); tensor_of_value(vvt, &vt); tensor_of_value(vin, &a);
  nxc_status s = nxc_svd(CTX_OF(vu), &u, &sv, &vt, &a);
  if (s) raise_status("svd", CTX_OF(vu), s);
  CAMLreturn(Val_unit);
}
/* both eig and eigvals raise as "eig", like the reference stub (nx_c_eig.c:1310-1326) */
CAMLprim value nx_cuda_eig(value vw, value vv, value vin, value vvectors) {
  CAMLparam4(vw, vv, vin, vvectors);
  nxc_tensor w, v, a; tensor_of_value(vw, &w); tensor_of_value(vv, &v); tensor_of_value(vin, &a);
  nxc_status s = nxc_eig(CTX_OF(vw), &w, &v, &a, Bool_val(vvectors));
  if (s) raise_status("eig", CTX_OF(vw), s);
  CAMLreturn(Val_unit);
}

/* ---- step capture (nxc_capture_begin / _end / nxc_graph_launch; no reference counterpart: the
   reference's answer to per-op overhead is Rune.jit, an eager device backend's is replay) ---- */
typedef struct { nxc_graph *g; nxc_ctx *ctx; } graphbox;
#define Graph_val(v) ((graphbox *)Data_custom_val(v))
static void graph_finalize(value v) {
  graphbox *b = Graph_val(v);
  if (b->g) nxc_graph_destroy(b->ctx, b->g);
}
static struct custom_operations graph_ops = {"nx_cuda.graph", graph_finalize, custom_compare_default,
    custom_hash_default, custom_serialize_default, custom_deserialize_default, custom_compare_ext_default,
    custom_fixed_length_default};
CAMLprim value nx_cuda_capture_begin(value vctx) {
  CAMLparam1(vctx);
  nxc_status s = nxc_capture_begin(Ctx_val(vctx));
  if (s) raise_status("capture_begin", Ctx_val(vctx), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_capture_end(value vctx) {
  CAMLparam1(vctx);
  CAMLlocal1(v);
  v = caml_alloc_custom(&graph_ops, sizeof(graphbox), 0, 1);
  Graph_val(v)->g = NULL;
  Graph_val(v)->ctx = Ctx_val(vctx);
  nxc_graph *g = NULL;
  nxc_status s = nxc_capture_end(Ctx_val(vctx), &g);
  if (s) raise_status("capture_end", Ctx_val(vctx), s);
  Graph_val(v)->g = g;
  CAMLreturn(v);
}
CAMLprim value nx_cuda_graph_launch(value vctx, value vg) {
  CAMLparam2(vctx, vg);
  nxc_status s = nxc_graph_launch(Ctx_val(vctx), Graph_val(vg)->g);
  if (s) raise_status("graph_launch", Ctx_val(vctx), s);
  CAMLreturn(Val_unit);
}
CAMLprim value nx_cuda_sync(value vctx) {
  CAMLparam1(vctx);
  nxc_ctx *ctx = Ctx_val(vctx);
  caml_enter_blocking_section();
  nxc_status s = nxc_sync(ctx);
  caml_leave_blocking_section();
  if (s) raise_status("sync", ctx, s);
  CAMLreturn(Val_unit);
}
