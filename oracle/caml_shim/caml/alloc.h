/* oracle/caml_shim/caml/alloc.h -- TEST INFRASTRUCTURE ONLY (see mlvalues.h).
 * Included by the reference's buffer header; nothing from it is used on the
 * path the oracle exercises. */
#ifndef NXREF_CAML_ALLOC_H
#define NXREF_CAML_ALLOC_H
#include "mlvalues.h"
#endif
