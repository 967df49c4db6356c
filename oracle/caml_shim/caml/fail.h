/* oracle/caml_shim/caml/fail.h -- TEST INFRASTRUCTURE ONLY (see mlvalues.h).
 * The two raisers are implemented in oracle/ref_runtime.c: they record the
 * message + exception class and longjmp back to nxref_invoke. */
#ifndef NXREF_CAML_FAIL_H
#define NXREF_CAML_FAIL_H
#include "mlvalues.h"
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#endif
