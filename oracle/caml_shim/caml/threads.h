/* oracle/caml_shim/caml/threads.h -- TEST INFRASTRUCTURE ONLY (see mlvalues.h).
 * There is no OCaml runtime lock in the harness; both are no-ops
 * (oracle/ref_runtime.c). */
#ifndef NXREF_CAML_THREADS_H
#define NXREF_CAML_THREADS_H
void caml_enter_blocking_section(void);
void caml_leave_blocking_section(void);
#define caml_release_runtime_system caml_enter_blocking_section
#define caml_acquire_runtime_system caml_leave_blocking_section
#endif
