/* tests/ocaml_api/caml/bigarray.h -- TEST INFRASTRUCTURE ONLY (see custom.h here): the oracle shim's
 * struct caml_ba_array plus the documented size query. */
#ifndef NXTEST_CAML_BIGARRAY_H
#define NXTEST_CAML_BIGARRAY_H
#include "../../../oracle/caml_shim/caml/bigarray.h"
uintnat caml_ba_byte_size(struct caml_ba_array *b);
#endif
