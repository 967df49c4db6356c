/* oracle/ref_runtime.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The ~40 lines of "OCaml runtime" the reference's nx_c_*.c need when they are
 * compiled (unmodified, from /root/reference) into oracle/_ref/libnxref.so:
 * the two exception raisers and the runtime-lock handshake. A raise records the
 * message and exception class and longjmps back to nxref_invoke, which is the
 * only way the Python harness (oracle/ref.py) enters a caml_nx_c_* stub.
 *
 * Nothing here is shipped: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the resulting library.
 */
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <caml/mlvalues.h>

static __thread jmp_buf nxref_jmp;
static __thread int nxref_armed = 0;
static __thread int nxref_class = 0; /* 1 = Failure, 2 = Invalid_argument */
static __thread char nxref_msg[512];

static void nxref_raise(int cls, const char *msg) __attribute__((noreturn));
static void nxref_raise(int cls, const char *msg) {
  nxref_class = cls;
  snprintf(nxref_msg, sizeof nxref_msg, "%s", msg ? msg : "");
  if (!nxref_armed) {
    fprintf(stderr, "nxref: exception outside nxref_invoke: %s\n", nxref_msg);
    __builtin_trap();
  }
  longjmp(nxref_jmp, 1);
}

void caml_failwith(const char *msg) { nxref_raise(1, msg); }
void caml_invalid_argument(const char *msg) { nxref_raise(2, msg); }
void caml_enter_blocking_section(void) {}
void caml_leave_blocking_section(void) {}

const char *nxref_last_message(void) { return nxref_msg; }

typedef value (*nxref_fn1)(value);
typedef value (*nxref_fn2)(value, value);
typedef value (*nxref_fn3)(value, value, value);
typedef value (*nxref_fn4)(value, value, value, value);
typedef value (*nxref_fn5)(value, value, value, value, value);
typedef value (*nxref_fn6)(value, value, value, value, value, value);
typedef value (*nxref_fn7)(value, value, value, value, value, value, value);

/* Call a CAMLprim stub with `nargs` value arguments. Returns 0 on success,
   1 if it raised Failure, 2 if it raised Invalid_argument. */
int nxref_invoke(void *fn, int nargs, const value *a) {
  nxref_class = 0;
  nxref_msg[0] = 0;
  nxref_armed = 1;
  if (setjmp(nxref_jmp) != 0) {
    nxref_armed = 0;
    return nxref_class;
  }
  switch (nargs) {
    case 1: ((nxref_fn1)fn)(a[0]); break;
    case 2: ((nxref_fn2)fn)(a[0], a[1]); break;
    case 3: ((nxref_fn3)fn)(a[0], a[1], a[2]); break;
    case 4: ((nxref_fn4)fn)(a[0], a[1], a[2], a[3]); break;
    case 5: ((nxref_fn5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: ((nxref_fn6)fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 7: ((nxref_fn7)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
    default: nxref_armed = 0; return -1;
  }
  nxref_armed = 0;
  return 0;
}
