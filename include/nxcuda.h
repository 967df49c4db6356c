/* nxcuda.h -- the C ABI of libnxcuda, the B200-native engine behind Nx's
 * `nx.backend` seam.
 *
 * This is the drop-in boundary. An OCaml library `packages/nx-cuda`
 * ((implements nx.backend)) binds these entry points 1:1 through OCaml
 * externals, exactly as the reference's veneer binds its own C stubs
 * (reference: packages/nx/lib/backend_c/nx_backend.ml:104-168, 242-264, 349,
 * 482-483). Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions (all mirrored from the reference so the OCaml side is plumbing):
 *   - dtype tags are Dtype.Packed.tag (reference: nx_c.h:130-168, 197-200).
 *   - a tensor operand is {buffer, shape, strides, offset}, strides and offset
 *     in ELEMENTS; strides may be 0 (broadcast) or negative (flip)
 *     (reference: nx_c.h:47-61, 406-412).
 *   - the OUTPUT is the first tensor argument and is allocated by the caller
 *     (reference: nx_c_engine.h:256-262).
 *   - a status is NULL on success, else a static never-freed string with the
 *     same text as the reference's (reference: nx_c.h:378-388,
 *     nx_c_engine.h:38-50). The binding raises "<op>: <status>" as
 *     Invalid_argument when nxc_status_is_invalid_argument() says so, Failure
 *     otherwise (reference: nx_c_engine.c:1345-1351).
 *   - every launch is asynchronous on the context's stream; only nxc_sync and
 *     nxc_d2h block (nxc_sync also drains the copy / communication side streams). There is no CPU fallback anywhere in this library.
 */
#ifndef NXCUDA_H
#define NXCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NXC_API __attribute__((visibility("default")))
#else
#define NXC_API
#endif

#define NXC_MAX_NDIM 32 /* reference: nx_c.h:45 */

typedef const char *nxc_status; /* reference: nx_c.h:378-379 */
#define NXC_OK ((nxc_status)0)

/* Dtype.Packed.tag order (reference: nx_c.h:130-168). */
typedef enum {
  NXC_F16 = 0, NXC_F32 = 1, NXC_F64 = 2, NXC_BF16 = 3, NXC_F8E4M3 = 4,
  NXC_F8E5M2 = 5, NXC_I4 = 6, NXC_U4 = 7, NXC_I8 = 8, NXC_U8 = 9, NXC_I16 = 10,
  NXC_U16 = 11, NXC_I32 = 12, NXC_U32 = 13, NXC_I64 = 14, NXC_U64 = 15,
  NXC_C32 = 16, NXC_C64 = 17, NXC_BOOL = 18, NXC_DTYPE_COUNT = 19
} nxc_dtype;

/* Operand descriptor; replaces the slots 0-3 the reference reads out of the
   OCaml record (reference: nx_c.h:406-434). `data` is a DEVICE pointer to
   element 0 of the buffer; the first live element is data + offset*elem_size. */
typedef struct {
  void *data;
  int32_t dtype;
  int32_t ndim;
  int64_t shape[NXC_MAX_NDIM];
  int64_t strides[NXC_MAX_NDIM];
  int64_t offset;
} nxc_tensor;

/* Unary ops, in the reference's stub order (reference: nx_c_map.c:1184-1204). */
typedef enum {
  NXC_NEG = 0, NXC_RECIP, NXC_ABS, NXC_SIGN, NXC_SQRT, NXC_EXP, NXC_LOG, NXC_SIN,
  NXC_COS, NXC_TAN, NXC_ASIN, NXC_ACOS, NXC_ATAN, NXC_SINH, NXC_COSH, NXC_TANH,
  NXC_TRUNC, NXC_CEIL, NXC_FLOOR, NXC_ROUND, NXC_ERF, NXC_MAP1_COUNT
} nxc_map1_op;

/* Binary ops (reference: nx_c_map.c:1207-1221). */
typedef enum {
  NXC_ADD = 0, NXC_SUB, NXC_MUL, NXC_IDIV, NXC_FDIV, NXC_MOD, NXC_MAX, NXC_MIN,
  NXC_POW, NXC_ATAN2, NXC_XOR, NXC_OR, NXC_AND, NXC_SHL, NXC_SHR, NXC_MAP2_COUNT
} nxc_map2_op;

/* Comparisons, bool output (reference: nx_c_map.c:1253-1256). */
typedef enum { NXC_CMPEQ = 0, NXC_CMPNE, NXC_CMPLT, NXC_CMPLE, NXC_CMP_COUNT } nxc_cmp_op;

/* Reductions / scans (reference: nx_c_fold.c:824-837). */
typedef enum { NXC_SUM = 0, NXC_PROD, NXC_RMAX, NXC_RMIN, NXC_REDUCE_COUNT } nxc_reduce_op;

typedef struct nxc_ctx nxc_ctx;

/* ---- context, memory, transfer ------------------------------------------
   replaces: `create_context : unit -> context` (reference:
   backend/nx_backend.mli:32-39) and Nx_buffer.create / to_host / from_host
   (reference: backend_c/nx_backend.ml:50-69). Device index comes from
   NX_CUDA_DEVICE (default: LOCAL_RANK, else 0); matmul precision from
   NX_CUDA_MATMUL in {f32 (default), ieee, tf32}: see nxc_set_matmul_mode. */
NXC_API nxc_status nxc_ctx_create(nxc_ctx **out);
/* Same, on an explicit device and an existing cudaStream_t (NULL = own stream). */
NXC_API nxc_status nxc_ctx_create_on(int device, void *cuda_stream, nxc_ctx **out);
NXC_API void nxc_ctx_destroy(nxc_ctx *ctx);
NXC_API nxc_status nxc_sync(nxc_ctx *ctx);
NXC_API void *nxc_stream(nxc_ctx *ctx); /* the cudaStream_t launches go to */
NXC_API int nxc_device(nxc_ctx *ctx);
/* Detail text of the most recent non-OK status on this context (CUDA error
   string included); never NULL. */
NXC_API const char *nxc_last_error(nxc_ctx *ctx);
/* Kernels launched by this library on this context since creation. */
NXC_API uint64_t nxc_launch_count(nxc_ctx *ctx);
/* 1 -> Invalid_argument, 0 -> Failure (reference: nx_c_engine.c:1345-1351,
   nx_c_matmul.c:1229-1237). */
NXC_API int nxc_status_is_invalid_argument(nxc_status s);
NXC_API int64_t nxc_elem_size(int dtype); /* 0 for packed int4/uint4 (reference: nx_c.h:309-321) */
/* How f32 operands are multiplied; 0 ok, -1 unknown mode. "f32" (default): f32-class accuracy --
   products of a few GFLOP and more run on the tensor cores as 3xTF32 (hi / lo split, error inside
   the classical K*u sgemm bound), smaller ones on the CUDA-core kernel; "ieee": always the
   CUDA-core kernel (every product and sum rounded to nearest); "tf32": plain tf32 operands,
   1e-3 relative, opt-in. bf16 / f16 always use the tensor cores with f32 accumulation. */
NXC_API int nxc_set_matmul_mode(nxc_ctx *ctx, const char *mode);

NXC_API nxc_status nxc_alloc(nxc_ctx *ctx, size_t bytes, void **dptr); /* stream-ordered, cached */
NXC_API nxc_status nxc_free(nxc_ctx *ctx, void *dptr);                 /* ordered after in-flight work */
NXC_API nxc_status nxc_host_alloc(nxc_ctx *ctx, size_t bytes, void **hptr); /* pinned */
NXC_API nxc_status nxc_host_free(nxc_ctx *ctx, void *hptr);
NXC_API nxc_status nxc_h2d(nxc_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async */
NXC_API nxc_status nxc_d2h(nxc_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* blocks */
NXC_API nxc_status nxc_memset(nxc_ctx *ctx, void *dst_dev, int byte, size_t bytes);
/* Copy engines. nxc_h2d from PINNED host memory (nxc_host_alloc) of >= 1 MiB runs on a dedicated
   host->device stream, ordered after the work already queued and before the work queued later on
   the context stream. nxc_d2h_async copies into PINNED host memory on a dedicated device->host
   stream, ordered after the work already queued; it does not block and later kernels do not wait
   for it, so step i's read-back overlaps step i+1's upload and compute. The data is in dst_host
   after nxc_sync (which drains every stream of the context). nxc_free of src_dev may be called
   at once: the engine defers the release until the copy has finished. */
NXC_API nxc_status nxc_d2h_async(nxc_ctx *ctx, void *dst_host_pinned, const void *src_dev, size_t bytes);

/* ---- step capture (new; no reference counterpart -- the reference's eager backend pays one
   function call per op, a device backend pays one LAUNCH per op, and a training step is a few
   thousand of them). Between nxc_capture_begin and nxc_capture_end every call on the context is
   recorded into a CUDA graph instead of running; nxc_graph_launch replays the whole step with one
   launch, any number of times, on the context's stream.
     - memory: while capturing, nxc_alloc / nxc_free are served by an arena the graph owns, so
       the handles created during the capture keep their addresses and are the replay's outputs;
       nxc_free of one outside the capture only drops the handle. The arena is released when the
       graph has been destroyed AND the last handle into it has been dropped, in either order (a
       garbage collector finalises in no particular order).
       Buffers that existed before the capture (inputs, parameters) are read in place by every
       replay: keep them alive and update them in place (nxc_copy into them) between replays.
     - calls that block or read back (nxc_sync, nxc_d2h, the checked gather / scatter range test,
       the linalg status words) return "operation not allowed while a step is being captured";
       nxc_gather / nxc_scatter defer their range check to the next nxc_sync as with
       NX_CUDA_SYNC_CHECKS=0. nxc_h2d / nxc_d2h_async are recorded as copy nodes and need
       page-locked host memory (a replay re-reads / re-writes the same host buffer).
     - collectives are capturable (the peer-memory exchanges keep their epoch on the device);
       every rank must replay the same graphs in the same order.
   nxc_graph_kernels = kernel nodes in the graph (what one replay adds to nxc_launch_count). */
typedef struct nxc_graph nxc_graph;
NXC_API nxc_status nxc_capture_begin(nxc_ctx *ctx);
NXC_API nxc_status nxc_capture_end(nxc_ctx *ctx, nxc_graph **out);
NXC_API nxc_status nxc_graph_launch(nxc_ctx *ctx, nxc_graph *g);
NXC_API uint64_t nxc_graph_kernels(nxc_graph *g);
NXC_API size_t nxc_graph_arena_bytes(nxc_graph *g);
NXC_API void nxc_graph_destroy(nxc_ctx *ctx, nxc_graph *g);

/* ---- map family -----------------------------------------------------------
   replaces caml_nx_c_{neg..erf}, caml_nx_c_{add..shr}, caml_nx_c_cmp*,
   caml_nx_c_where, caml_nx_c_cast, caml_nx_c_copy (reference:
   nx_c_map.c:1184-1280, nx_c_move.c:191-201) and Nx_buffer.fill as used by
   `full` (reference: backend_c/nx_backend.ml:61-64). All operands carry
   out->ndim dims of out->shape; input strides may be 0. */
NXC_API nxc_status nxc_map1(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *a);
NXC_API nxc_status nxc_map2(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *a,
                    const nxc_tensor *b);
NXC_API nxc_status nxc_cmp(nxc_ctx *ctx, int op, const nxc_tensor *out_bool, const nxc_tensor *a,
                   const nxc_tensor *b);
NXC_API nxc_status nxc_where(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *cond,
                     const nxc_tensor *a, const nxc_tensor *b);
NXC_API nxc_status nxc_cast(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a);
NXC_API nxc_status nxc_copy(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a);
/* `scalar` points at ONE element in out's storage type, on the HOST; it is
   passed as a kernel argument (no H2D copy, no sync). */
NXC_API nxc_status nxc_fill(nxc_ctx *ctx, const nxc_tensor *out, const void *scalar);

/* ---- fold family ----------------------------------------------------------
   replaces caml_nx_c_reduce_{sum,prod,max,min} and caml_nx_c_{argmax,argmin}
   (reference: nx_c_fold.c:824-833, nx_c_engine.c:1052-1254, 1392-1473).
   `axes` strictly increasing; `out` is either squeezed (rank in->ndim - n) or
   full rank with size-1 reduced dims. argreduce writes int32. */
NXC_API nxc_status nxc_reduce(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *in,
                      const int *axes, int n_axes);
NXC_API nxc_status nxc_argreduce(nxc_ctx *ctx, int is_max, const nxc_tensor *out_i32,
                         const nxc_tensor *in, int axis);
/* replaces caml_nx_c_cum{sum,prod,max,min} (reference: nx_c_fold.c:834-837). */
NXC_API nxc_status nxc_scan(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *in, int axis);

/* ---- fft family ------------------------------------------------------------
   replaces caml_nx_c_fft / caml_nx_c_ifft / caml_nx_c_rfft / caml_nx_c_irfft
   (reference: nx_c_fft.c:1173-1223; drivers :940-1143). UNNORMALISED transforms
   (fft = -sign DFT, ifft = +sign DFT without 1/n), double-precision arithmetic for
   both c32 and c64, any length. `out` is allocated by the binding: the input's
   shape for fft/ifft; last transformed axis n/2+1 for rfft; `s_last` (or
   2*(half-1) when s_last <= 0) for irfft, which truncates or zero-pads the
   half-spectrum. Axis out of range / shape mismatch are Invalid_argument;
   a non-complex (fft, irfft input) or non-real (rfft input) dtype is
   Failure "unsupported bigarray kind". */
NXC_API nxc_status nxc_fft(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const int *axes,
                           int n_axes, int inverse);
NXC_API nxc_status nxc_rfft(nxc_ctx *ctx, const nxc_tensor *out_complex, const nxc_tensor *in_real,
                            const int *axes, int n_axes);
NXC_API nxc_status nxc_irfft(nxc_ctx *ctx, const nxc_tensor *out_real, const nxc_tensor *in_complex,
                             const int *axes, int n_axes, int64_t s_last);

/* ---- linalg tier 1 ---------------------------------------------------------
   replaces caml_nx_c_cholesky, caml_nx_c_triangular_solve (reference:
   nx_c_tri.c:570-620) and caml_nx_c_qr (reference: nx_c_qr.c:421-433). Batched over
   the leading dims, operands at arbitrary strides, outputs allocated by the binding
   (cholesky / solve: the input / rhs shape; qr: reduced [m,k],[k,n] or full
   [m,m],[m,n]). f16/bf16/fp8/f32 compute in f32; f64, c32, c64 natively; other
   dtypes are Invalid_argument "linalg requires a float or complex dtype".
   `flags` of the solve: bit 0 upper, bit 1 (conjugate) transpose, bit 2 unit
   diagonal. Numeric failures are Failure "matrix is not positive definite" /
   "triangular matrix is singular" (the OCaml veneer lifts them to Linalg_error,
   backend_c/nx_backend.ml:582-594); the call reads one status word back, so these
   three synchronise like the reference's do. */
NXC_API nxc_status nxc_cholesky(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int upper);
NXC_API nxc_status nxc_triangular_solve(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a,
                                        const nxc_tensor *b, int flags);
NXC_API nxc_status nxc_qr(nxc_ctx *ctx, const nxc_tensor *q, const nxc_tensor *r, const nxc_tensor *in,
                          int reduced);
/* eigh / eigvalsh (linalg tier 2): replaces caml_nx_c_eigh (reference: nx_c_eigh.c; veneer
   backend_c/nx_backend.ml:627-648). Reads the LOWER triangle of the Hermitian input; `w_f64`
   ([batch..., n], always float64) receives the eigenvalues in ascending order; `v` (input dtype
   and shape, columns = eigenvectors) is written only when vectors != 0 (pass the input there
   otherwise, as the reference veneer does). Failure "eigenvalue iteration did not converge". */
NXC_API nxc_status nxc_eigh(nxc_ctx *ctx, const nxc_tensor *w_f64, const nxc_tensor *v, const nxc_tensor *in,
                            int vectors);
/* svd (linalg tier 3): replaces caml_nx_c_svd (reference: nx_c_svd.c:2943-2958, driver
   nx_c_svd_run :2743-2938; veneer backend_c/nx_backend.ml:650-677). in [batch..., m, n], k = min(m, n):
   `s_f64` [batch..., k] (always float64) receives the singular values, descending and non-negative;
   `u` [batch..., m, k|m] and `vt` [batch..., k|n, n] (input dtype) receive U and V^H -- thin or full is
   read off each output's shape, as the reference does. Failure "eigenvalue iteration did not
   converge" (the reference's LA_ERR_NO_CONVERGE, lifted to Linalg_error by the veneer). */
NXC_API nxc_status nxc_svd(nxc_ctx *ctx, const nxc_tensor *u, const nxc_tensor *s_f64, const nxc_tensor *vt,
                           const nxc_tensor *in);
/* eig / eigvals (linalg tier 3): replaces caml_nx_c_eig (reference: nx_c_eig.c:1310-1326, driver
   nx_c_eig_run :1195-1298; veneer backend_c/nx_backend.ml:679-707). General square matrices of any
   float / complex dtype; `w_c64` [batch..., n] and `v_c64` [batch..., n, n] are ALWAYS complex128;
   eigenvector columns have unit 2-norm, no phase or order convention; `v_c64` is written only when
   vectors != 0 (pass `w_c64` there otherwise, as the reference veneer does). Invalid_argument
   "eig requires a float or complex dtype"; Failure "eigenvalue iteration did not converge". */
NXC_API nxc_status nxc_eig(nxc_ctx *ctx, const nxc_tensor *w_c64, const nxc_tensor *v_c64, const nxc_tensor *in,
                           int vectors);

/* ---- matmul ---------------------------------------------------------------
   replaces caml_nx_c_matmul (reference: nx_c_matmul.c:874-1108, 1271-1277).
   A [...,m,k] and B [...,k,n] at arbitrary strides with broadcast batch dims;
   out [batch...,m,n]. bf16/f16 run on tcgen05 tensor cores with f32
   accumulation in TMEM; f32 is exact FFMA unless the context is in tf32 mode;
   f64/int/complex/fp8 run on CUDA cores in the reference's compute type. */
NXC_API nxc_status nxc_matmul(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a,
                      const nxc_tensor *b);

/* ---- move / index / random ("next" rows of the scope table) ----------------
   replaces caml_nx_c_pad, _cat, _gather, _scatter, _threefry (reference:
   nx_c_move.c:229-569, nx_c_random.c:44-140). */
NXC_API nxc_status nxc_pad(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in,
                   const void *fill_scalar_host, const int64_t *pad_before);
NXC_API nxc_status nxc_cat(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *const *ins, int n_in,
                   int axis);
NXC_API nxc_status nxc_gather(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data,
                      const nxc_tensor *indices_i32, int axis);
/* nxc_gather for indices the backend produced itself (argmax / argmin / argsort results): the
   range flag is not read back, so the call does not drain the stream. Out-of-range indices
   are still never dereferenced, and the flag is sticky: the next nxc_sync / nxc_d2h raises
   "index out of bounds for the gathered/scattered axis". */
NXC_API nxc_status nxc_gather_trusted(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data,
                                      const nxc_tensor *indices_i32, int axis);
/* `out` is pre-seeded with the template; mode 0 = Set (last write wins), 1 = Add. */
NXC_API nxc_status nxc_scatter(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *indices_i32,
                       const nxc_tensor *updates, int axis, int mode);
NXC_API nxc_status nxc_threefry(nxc_ctx *ctx, const nxc_tensor *out_i32, const nxc_tensor *key_i32,
                        const nxc_tensor *ctr_i32);

/* replaces caml_nx_c_sort / caml_nx_c_argsort (reference: nx_c_sort.c:476-514). NaN-class
   elements last in both directions; complex lexicographic; argsort stable, int32 result. */
NXC_API nxc_status nxc_sort(nxc_ctx *ctx, int is_argsort, const nxc_tensor *out, const nxc_tensor *in, int axis,
                            int descending);

/* replaces caml_nx_c_unfold / caml_nx_c_fold (reference: nx_c_move.c:588-870): im2col / col2im
   over the trailing K spatial dims; padding_flat = [before0, after0, before1, ...]. */
NXC_API nxc_status nxc_unfold(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int K,
                              const int64_t *kernel, const int64_t *stride, const int64_t *dilation,
                              const int64_t *padding_flat);
NXC_API nxc_status nxc_fold(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int K,
                            const int64_t *output_size, const int64_t *kernel, const int64_t *stride,
                            const int64_t *dilation, const int64_t *padding_flat);

/* ---- multi-GPU exchange (new; the reference has no collective in the
   backend contract -- SURVEY.md section 8e). One process per GPU; NCCL is
   dlopen'ed (libnccl.so.2) so the library loads without it. ------------------ */
#define NXC_UNIQUE_ID_BYTES 128
NXC_API nxc_status nxc_dist_unique_id(void *id_out_128);
NXC_API nxc_status nxc_dist_init(nxc_ctx *ctx, int rank, int world, const void *id_128);
NXC_API nxc_status nxc_dist_finalize(nxc_ctx *ctx);
/* 1 when nxc_dist_init mapped every peer's mailbox (CUDA IPC over NVLink): nxc_allreduce /
   nxc_allgather of payloads up to 256 KiB per rank then run as ONE kernel that stores into the
   peers' memory and waits on flags (nxc_allreduce and nxc_argreduce_exchange fold what they
   receive in the same kernel); larger payloads, the async entry point and NX_CUDA_P2P=0 use NCCL. */
NXC_API int nxc_dist_p2p_enabled(nxc_ctx *ctx);
/* In-place allreduce over `count` elements of `dtype`; op is an nxc_reduce_op. Every dtype / op the
   single-device nxc_reduce takes. Up to 256 KiB with the mailboxes mapped: ONE kernel pushes the
   partial to every peer, waits for theirs and folds them in rank order (bit-identical on all
   ranks, float max / min NaN-sticky). Larger: NCCL where it has the reduction, else all-gather +
   the backend's own reduce over the rank axis. A peer that never shows up is reported by the
   next nxc_sync / nxc_d2h as "NCCL error" (detail: the exchange timed out) -- never a hang, never
   silently undefined data. */
NXC_API nxc_status nxc_allreduce(nxc_ctx *ctx, void *dev_buf, int64_t count, int dtype, int op);
/* The finish of an argmax / argmin along an axis whose slabs live on the ranks, in ONE kernel over
   the mailboxes: `local_idx_i32` is this rank's nxc_argreduce result over its slab `x_local`
   along `axis` (squeezed, C-contiguous), `idx_offset` the global position of the slab's first
   element along that axis; `out_i32` (same shape, C-contiguous) receives the GLOBAL index of the
   extreme over all ranks by the single-device rule (first index on ties, first NaN wins;
   reference: nx_c_fold.c:93-101). At most nxc_argreduce_exchange_max_outputs outputs (0 when the
   mailboxes are not mapped: use all-gathers then). */
NXC_API nxc_status nxc_argreduce_exchange(nxc_ctx *ctx, int is_max, const nxc_tensor *out_i32,
                                          const nxc_tensor *x_local, const nxc_tensor *local_idx_i32, int axis,
                                          int64_t idx_offset);
NXC_API int64_t nxc_argreduce_exchange_max_outputs(nxc_ctx *ctx);
NXC_API nxc_status nxc_allgather(nxc_ctx *ctx, const void *dev_send, void *dev_recv,
                         int64_t bytes_per_rank);
/* The same allreduce on the context's communication stream: ordered after the work already
   queued, but kernels queued later do NOT wait for it (a gradient bucket reduces under the rest
   of the backward pass). nxc_comm_wait makes everything queued afterwards wait for all async
   collectives issued so far (no host block). The buffer must stay allocated until then. */
NXC_API nxc_status nxc_allreduce_async(nxc_ctx *ctx, void *dev_buf, int64_t count, int dtype, int op);
NXC_API nxc_status nxc_comm_wait(nxc_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NXCUDA_H */
