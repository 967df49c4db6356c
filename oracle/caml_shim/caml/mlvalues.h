/* oracle/caml_shim/caml/mlvalues.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for the OCaml runtime's <caml/mlvalues.h>, just enough for
 * the reference's nx_c_*.c translation units to compile WITHOUT an OCaml
 * toolchain (none exists in this image). It reproduces OCaml's public value
 * encoding (immediates are (n<<1)|1, a block is a pointer to field 0 with a
 * header word at index -1 whose size lives above bit 10), which is documented
 * OCaml ABI, not reference code. The Python harness (oracle/ref.py) builds
 * operand records in this layout and calls the reference's caml_nx_c_* stubs.
 */
#ifndef NXREF_CAML_MLVALUES_H
#define NXREF_CAML_MLVALUES_H
#include <stddef.h>
#include <stdint.h>

typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef intnat value;
typedef uintnat header_t;
typedef uintnat mlsize_t;

#define Is_long(x) (((x) & 1) != 0)
#define Is_block(x) (((x) & 1) == 0)
#define Val_long(x) ((value)(((uintnat)(intnat)(x) << 1) + 1))
#define Long_val(x) ((intnat)(x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int)Long_val(x))
#define Val_unit Val_long(0)
#define Val_bool(x) Val_long((x) != 0)
#define Bool_val(x) Int_val(x)
#define Val_false Val_long(0)
#define Val_true Val_long(1)

#define Field(v, i) (((value *)(v))[i])
#define Hd_val(v) (((header_t *)(v))[-1])
#define Wosize_hd(hd) ((mlsize_t)((hd) >> 10))
#define Wosize_val(v) (Wosize_hd(Hd_val(v)))

#define CAMLprim
#define CAMLextern extern
#define CAMLnoreturn_start
#define CAMLnoreturn_end __attribute__((noreturn))

#endif
