/* tests/ocaml_api/caml/custom.h -- TEST INFRASTRUCTURE ONLY.
 * Declarations of the PUBLIC OCaml C API for custom blocks (OCaml manual, "Interfacing C with
 * OCaml", section "Custom blocks": struct custom_operations and caml_alloc_custom[_mem]), written
 * from the documented interface so that packages/nx-cuda/lib/nx_cuda_stubs.c can be type-checked
 * in an image without an OCaml toolchain (tests/test_ocaml_stubs_compile.py). Layered over the
 * oracle's minimal shim (oracle/caml_shim), which holds the value encoding. */
#ifndef NXTEST_CAML_CUSTOM_H
#define NXTEST_CAML_CUSTOM_H
#include <stddef.h>
#include <caml/mlvalues.h>
struct custom_fixed_length { intnat bsize_32; intnat bsize_64; };
struct custom_operations {
  const char *identifier;
  void (*finalize)(value v);
  int (*compare)(value v1, value v2);
  intnat (*hash)(value v);
  void (*serialize)(value v, uintnat *bsize_32, uintnat *bsize_64);
  uintnat (*deserialize)(void *dst);
  int (*compare_ext)(value v1, value v2);
  const struct custom_fixed_length *fixed_length;
};
#define custom_finalize_default NULL
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL
#define custom_fixed_length_default NULL
#define Data_custom_val(v) ((void *)&Field((v), 1))
value caml_alloc_custom(struct custom_operations *ops, uintnat size, mlsize_t mem, mlsize_t max);
value caml_alloc_custom_mem(struct custom_operations *ops, uintnat size, mlsize_t mem);
#endif
