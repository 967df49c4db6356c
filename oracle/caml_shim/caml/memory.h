/* oracle/caml_shim/caml/memory.h -- TEST INFRASTRUCTURE ONLY (see mlvalues.h).
 * GC rooting macros are no-ops: nothing on the hot path allocates on an OCaml
 * heap, and the harness keeps every block alive for the duration of the call. */
#ifndef NXREF_CAML_MEMORY_H
#define NXREF_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam0() ((void)0)
#define CAMLparam1(a) ((void)(a))
#define CAMLparam2(a, b) ((void)(a), (void)(b))
#define CAMLparam3(a, b, c) ((void)(a), (void)(b), (void)(c))
#define CAMLparam4(a, b, c, d) ((void)(a), (void)(b), (void)(c), (void)(d))
#define CAMLparam5(a, b, c, d, e) \
  ((void)(a), (void)(b), (void)(c), (void)(d), (void)(e))
#define CAMLxparam1(a) ((void)(a))
#define CAMLxparam2(a, b) ((void)(a), (void)(b))
#define CAMLxparam3(a, b, c) ((void)(a), (void)(b), (void)(c))
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLreturn(x) return (x)
#define CAMLreturn0 return
#endif
