/* tests/ocaml_api/caml/memory.h -- TEST INFRASTRUCTURE ONLY (see custom.h here): the oracle shim's
 * rooting macros plus the documented caml_stat_* allocation functions. */
#ifndef NXTEST_CAML_MEMORY_H
#define NXTEST_CAML_MEMORY_H
#include <stddef.h>
#include "../../../oracle/caml_shim/caml/memory.h"
typedef size_t asize_t;
void *caml_stat_alloc(asize_t size);
void caml_stat_free(void *p);
#endif
