"""packages/nx-cuda cannot be built here (no OCaml toolchain), but its C half can be type-checked:
the stubs compile warning-free against declarations of the public OCaml C API (tests/ocaml_api over
the oracle's value-encoding shim) and include/nxcuda.h, and every `external` of the OCaml veneer
names a CAMLprim the stubs define, with the same number of arguments."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "packages", "nx-cuda", "lib", "nx_cuda_stubs.c")
VENEER = os.path.join(ROOT, "packages", "nx-cuda", "lib", "nx_backend.ml")


def test_stubs_compile_against_the_ocaml_c_api_declarations():
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "tests", "ocaml_api"), "-I", os.path.join(ROOT, "oracle", "caml_shim"),
                        "-I", os.path.join(ROOT, "include"), STUBS], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def _prims():
    # after preprocessing, so that macro-generated stubs (STUB2(...)) are seen
    src = subprocess.run(["gcc", "-std=c11", "-E", "-P", "-I", os.path.join(ROOT, "tests", "ocaml_api"),
                          "-I", os.path.join(ROOT, "oracle", "caml_shim"), "-I", os.path.join(ROOT, "include"), STUBS],
                         capture_output=True, text=True, check=True).stdout
    out = {}
    for m in re.finditer(r"\bvalue\s+(nx_cuda_\w+)\s*\(([^)]*)\)\s*\{", src):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = len(args)
    return out


def _externals():
    src = open(VENEER).read()
    src = re.sub(r"\(\*.*?\*\)", " ", src, flags=re.S)
    out = []
    for m in re.finditer(r"external\s+(\w+)\s*:(.*?)=\s*((?:\"\w+\"\s*)+)", src, flags=re.S):
        ty, names = m.group(2), re.findall(r"\"(\w+)\"", m.group(3))
        depth, arrows, i = 0, 0, 0
        while i < len(ty):   # top-level arrows only: arguments may themselves be parenthesised types
            c = ty[i]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            elif c == "-" and ty[i:i + 2] == "->" and depth == 0:
                arrows += 1
            i += 1
        out.append((m.group(1), arrows, names))
    return out


def test_every_external_has_a_stub_of_the_same_arity():
    prims, ext = _prims(), _externals()
    assert len(ext) >= 30 and len(prims) >= 30
    for name, nargs, cnames in ext:
        if nargs <= 5:
            assert len(cnames) == 1, f"{name}: one C name expected for {nargs} arguments"
            assert cnames[0] in prims, f"{name}: stub {cnames[0]} is not defined"
            assert prims[cnames[0]] == nargs, f"{name}: {cnames[0]} takes {prims[cnames[0]]} values, external passes {nargs}"
        else:   # > 5 arguments: bytecode stub (argv, argn) + native stub (all arguments)
            assert len(cnames) == 2, f"{name}: {nargs} arguments need a bytecode and a native stub"
            assert cnames[0] in prims and prims[cnames[0]] == 2, f"{name}: bytecode stub {cnames[0]}"
            assert cnames[1] in prims and prims[cnames[1]] == nargs, f"{name}: native stub {cnames[1]}"
