/* tests/ocaml_rt/runtime_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The slice of the OCaml runtime that packages/nx-cuda/lib/nx_cuda_stubs.c links against, written
 * from the documented C interface ("Interfacing C with OCaml": custom blocks, caml_stat_*, the two
 * standard raisers, the blocking-section pair), so that the stubs can be EXECUTED in an image
 * without an OCaml toolchain -- the same trick oracle/ref_runtime.c plays for the reference's
 * stubs. A custom block is [header | ops pointer | data...]; a raise records the class and text
 * and longjmps back to nxstub_invoke, which is how the Python harness (tests/stubs_harness.py)
 * enters a stub. There is no GC: blocks live until nxstub_release runs their finalizer, which is
 * what the harness does when the Python wrapper is collected. */
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>

static __thread jmp_buf nxstub_jmp;
static __thread int nxstub_armed = 0;
static __thread int nxstub_class = 0; /* 1 = Failure, 2 = Invalid_argument */
static __thread char nxstub_msg[1024];
static long nxstub_live_blocks = 0;

static void nxstub_raise(int cls, const char *msg) __attribute__((noreturn));
static void nxstub_raise(int cls, const char *msg) {
  nxstub_class = cls;
  snprintf(nxstub_msg, sizeof nxstub_msg, "%s", msg ? msg : "");
  if (!nxstub_armed) {
    fprintf(stderr, "nxstub: exception outside nxstub_invoke: %s\n", nxstub_msg);
    abort();
  }
  longjmp(nxstub_jmp, 1);
}
void caml_failwith(const char *msg) { nxstub_raise(1, msg); }
void caml_invalid_argument(const char *msg) { nxstub_raise(2, msg); }
void caml_enter_blocking_section(void) {}
void caml_leave_blocking_section(void) {}
void *caml_stat_alloc(asize_t size) { return malloc(size ? size : 1); }
void caml_stat_free(void *p) { free(p); }

static const int ba_elem_size[] = {4, 8, 1, 1, 2, 2, 4, 8, 8, 8, 8, 16, 1, 2};
uintnat caml_ba_byte_size(struct caml_ba_array *b) {
  uintnat n = 1;
  for (intnat i = 0; i < b->num_dims; i++) n *= (uintnat)b->dim[i];
  return n * (uintnat)ba_elem_size[b->flags & CAML_BA_KIND_MASK];
}

value caml_alloc_custom(struct custom_operations *ops, uintnat size, mlsize_t mem, mlsize_t max) {
  (void)mem; (void)max;
  const uintnat words = 1 + (size + sizeof(value) - 1) / sizeof(value);
  header_t *blk = (header_t *)calloc(words + 1, sizeof(value));
  if (!blk) caml_failwith("out of memory");
  blk[0] = (header_t)words << 10 | 255; /* Custom_tag */
  value v = (value)(blk + 1);
  Field(v, 0) = (value)ops;
  nxstub_live_blocks++;
  return v;
}
value caml_alloc_custom_mem(struct custom_operations *ops, uintnat size, mlsize_t mem) {
  return caml_alloc_custom(ops, size, mem, 0);
}

/* what the GC would do to an unreachable custom block */
void nxstub_release(value v) {
  struct custom_operations *ops = (struct custom_operations *)Field(v, 0);
  if (ops && ops->finalize) ops->finalize(v);
  free((header_t *)v - 1);
  nxstub_live_blocks--;
}
long nxstub_live(void) { return nxstub_live_blocks; }
const char *nxstub_last_message(void) { return nxstub_msg; }
const char *nxstub_custom_identifier(value v) { return ((struct custom_operations *)Field(v, 0))->identifier; }

typedef value (*fn1)(value);
typedef value (*fn2)(value, value);
typedef value (*fn3)(value, value, value);
typedef value (*fn4)(value, value, value, value);
typedef value (*fn5)(value, value, value, value, value);
typedef value (*fn6)(value, value, value, value, value, value);
typedef value (*fn7)(value, value, value, value, value, value, value);
typedef value (*fnbc)(value *, int);

/* Call a stub with `nargs` value arguments (nargs < 0: the bytecode convention, (argv, -nargs)).
   Returns 0 and the result in *out, 1 if it raised Failure, 2 if it raised Invalid_argument. */
int nxstub_invoke(void *fn, int nargs, value *a, value *out) {
  nxstub_class = 0;
  nxstub_msg[0] = 0;
  nxstub_armed = 1;
  if (setjmp(nxstub_jmp) != 0) {
    nxstub_armed = 0;
    return nxstub_class;
  }
  value r = Val_unit;
  switch (nargs) {
    case 1: r = ((fn1)fn)(a[0]); break;
    case 2: r = ((fn2)fn)(a[0], a[1]); break;
    case 3: r = ((fn3)fn)(a[0], a[1], a[2]); break;
    case 4: r = ((fn4)fn)(a[0], a[1], a[2], a[3]); break;
    case 5: r = ((fn5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = ((fn6)fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 7: r = ((fn7)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
    default:
      if (nargs < 0) { r = ((fnbc)fn)(a, -nargs); break; }
      nxstub_armed = 0;
      return -1;
  }
  nxstub_armed = 0;
  if (out) *out = r;
  return 0;
}
