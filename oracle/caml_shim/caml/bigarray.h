/* oracle/caml_shim/caml/bigarray.h -- TEST INFRASTRUCTURE ONLY (see mlvalues.h).
 * Public layout of an OCaml Bigarray custom block (OCaml >= 5.2 kind order,
 * which added FLOAT16 = 13). Only `data` and `flags` are read on the path the
 * oracle exercises. */
#ifndef NXREF_CAML_BIGARRAY_H
#define NXREF_CAML_BIGARRAY_H
#include "mlvalues.h"

enum caml_ba_kind {
  CAML_BA_FLOAT32 = 0,
  CAML_BA_FLOAT64,
  CAML_BA_SINT8,
  CAML_BA_UINT8,
  CAML_BA_SINT16,
  CAML_BA_UINT16,
  CAML_BA_INT32,
  CAML_BA_INT64,
  CAML_BA_CAML_INT,
  CAML_BA_NATIVE_INT,
  CAML_BA_COMPLEX32,
  CAML_BA_COMPLEX64,
  CAML_BA_CHAR,
  CAML_BA_FLOAT16,
  CAML_BA_FIRST_UNIMPLEMENTED_KIND,
  CAML_BA_KIND_MASK = 0xFF
};

enum caml_ba_layout {
  CAML_BA_C_LAYOUT = 0,
  CAML_BA_FORTRAN_LAYOUT = 0x100,
  CAML_BA_LAYOUT_MASK = 0x100
};

struct caml_ba_proxy;

struct caml_ba_array {
  void *data;
  intnat num_dims;
  intnat flags;
  struct caml_ba_proxy *proxy;
  intnat dim[];
};

/* A bigarray value is a custom block: word 0 = ops pointer, struct follows. */
#define Caml_ba_array_val(v) ((struct caml_ba_array *)&Field(v, 1))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#endif
