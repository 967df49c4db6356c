"""Strided-view algebra for the host mirror: expand / reshape / permute / shrink /
flip rewrite (shape, strides, offset) and never touch data.

Behaviour follows the reference's View module (packages/nx/lib/core/view.ml:82-322)
and Shape.c_contiguous_strides (core/shape.ml:22-33): same results, same error
class (ValueError here stands for OCaml's Invalid_argument) and the same
"call contiguous() first" failure for a reshape the strides cannot express.
Strides and offset are in ELEMENTS.
"""
from __future__ import annotations


def c_contiguous_strides(shape):
    n = len(shape)
    st = [0] * n
    if n == 0:
        return st
    st[n - 1] = 0 if shape[n - 1] == 0 else 1
    for i in range(n - 2, -1, -1):
        st[i] = 0 if shape[i] == 0 else st[i + 1] * max(1, shape[i + 1])
    return st


def numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class View:
    __slots__ = ("shape", "strides", "offset")

    def __init__(self, shape, strides=None, offset=0):
        shape = tuple(int(s) for s in shape)
        zero = any(s == 0 for s in shape)
        if strides is None:
            strides = c_contiguous_strides(shape)
        elif len(strides) != len(shape):
            raise ValueError(f"create: strides length {len(strides)} != shape length {len(shape)}")
        self.shape = shape
        self.strides = tuple(int(s) for s in strides)
        self.offset = 0 if zero else int(offset)

    @property
    def ndim(self):
        return len(self.shape)

    def numel(self):
        return numel(self.shape)

    def is_c_contiguous(self):
        return list(self.strides) == c_contiguous_strides(self.shape)

    # -- movement ---------------------------------------------------------
    def expand(self, new_shape):
        new_shape = tuple(int(s) for s in new_shape)
        if self.ndim == 0:
            return View(new_shape, [0] * len(new_shape), self.offset)
        if len(new_shape) != self.ndim:
            raise ValueError(f"expand: rank mismatch: {len(new_shape)} vs {self.ndim}")
        if any(s == 0 for s in self.shape):
            return View(new_shape)
        st = []
        for i, ns in enumerate(new_shape):
            s = self.shape[i]
            if s == ns:
                st.append(self.strides[i])
            elif s == 1:
                st.append(0)
            else:
                raise ValueError(f"expand: dimension {i} (size {s}) cannot expand to size {ns}, "
                                 "only singletons expand")
        return View(new_shape, st, self.offset)

    def permute(self, axes):
        n = self.ndim
        if len(axes) != n:
            raise ValueError(f"permute: axes length {len(axes)} != ndim {n}")
        seen = set()
        for ax in axes:
            if ax < 0 or ax >= n:
                raise ValueError(f"permute: axis {ax} out of bounds for {n}D tensor")
            if ax in seen:
                raise ValueError(f"permute: duplicate axis {ax}")
            seen.add(ax)
        return View([self.shape[a] for a in axes], [self.strides[a] for a in axes], self.offset)

    def shrink(self, bounds):
        if len(bounds) != self.ndim:
            raise ValueError(f"shrink: bounds length {len(bounds)} != ndim {self.ndim}")
        if all(b == 0 and e == s for (b, e), s in zip(bounds, self.shape)):
            return self
        for (b, e), s in zip(bounds, self.shape):
            if b < 0 or e < 0 or b > s or e > s or b >= e:
                raise ValueError("shrink: bounds must be within shape and start < end")
        off = self.offset + sum(b * st for (b, _), st in zip(bounds, self.strides))
        return View([e - b for b, e in bounds], self.strides, off)

    def flip(self, flags):
        if len(flags) != self.ndim:
            raise ValueError(f"flip: boolean array length {len(flags)} != ndim {self.ndim}")
        off = self.offset
        st = list(self.strides)
        for i, f in enumerate(flags):
            if f and self.shape[i] > 0:
                off += (self.shape[i] - 1) * st[i]
                st[i] = -st[i]
        return View(self.shape, st, off)

    def reshape(self, new_shape):
        new_shape = tuple(int(s) for s in new_shape)
        if new_shape == self.shape:
            return self
        old_n, new_n = numel(self.shape), numel(new_shape)
        if old_n != new_n and old_n != 0 and new_n != 0:
            raise ValueError(f"reshape: cannot reshape {list(self.shape)} to {list(new_shape)}")
        if 0 in self.shape or 0 in new_shape:
            return View(new_shape)
        if self.is_c_contiguous():
            return View(new_shape, None, self.offset)
        if len(new_shape) == 0:
            return View(new_shape, None, self.offset)
        if all(s == 0 for s in self.strides):
            return View(new_shape, [0] * len(new_shape), self.offset)
        # only size-1 dims inserted / removed
        old_dims = [(s, st) for s, st in zip(self.shape, self.strides) if s != 1]
        new_core = [s for s in new_shape if s != 1]
        if [s for s, _ in old_dims] == new_core:
            it = iter(old_dims)
            st = [0 if s == 1 else next(it)[1] for s in new_shape]
            return View(new_shape, st, self.offset)
        # split / merge dims whose strides compose
        mapped = self._match(old_dims, new_core)
        if mapped is None:
            raise ValueError(
                f"reshape: cannot reshape {list(self.shape)} to {list(new_shape)}, incompatible strides "
                f"{list(self.strides)} (expected {c_contiguous_strides(new_shape)}), call contiguous() first")
        it = iter(mapped)
        st = [0 if s == 1 else next(it) for s in new_shape]
        return View(new_shape, st, self.offset)

    @staticmethod
    def _match(old, new):
        """Walk old (size, stride) dims and new sizes left to right, splitting an
        old dim over several new ones or merging several composing old dims into
        one new dim. Returns the new strides or None."""
        old = list(old)
        out = []
        i = 0
        for want in new:
            if i >= len(old):
                return None
            size, stride = old[i]
            if size == want:
                out.append(stride)
                i += 1
            elif size > want and size % want == 0:
                rest = size // want
                out.append(stride * rest)
                old[i] = (rest, stride)
            elif want > size:
                acc, st = size, stride
                j = i + 1
                while acc < want:
                    if j >= len(old):
                        return None
                    nsz, nst = old[j]
                    if st != nst * nsz:
                        return None
                    acc *= nsz
                    st = nst
                    j += 1
                if acc != want:
                    return None
                out.append(st)
                i = j
            else:
                return None
        return out if i == len(old) else None

    def __repr__(self):
        return f"View(shape={self.shape}, strides={self.strides}, offset={self.offset})"
