"""NTT / iNTT / coset LDE / NTT-based polynomial arithmetic: the mirror of src/fft/ntt.rs and
src/fft/ntt_arithmetics.rs (same names, argument order and failure behaviour; bodies run on
the B200 through libzkb200.so)."""
import ctypes

import numpy as np

from .context import Vec, default_context, le16, unpack


def _next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def ntt(root, inputs, ctx=None):
    """ntt(root, inputs) src/fft/ntt.rs:7-49: zero-pads to the next power of two."""
    ctx = ctx or default_context()
    v = Vec(inputs)
    out, optr = ctx.out_like(v, _next_pow2(v.n) if v.n else 1)
    ctx.check(ctx.lib.zkb_ntt(ctx.h, le16(root), v.ptr, v.n, optr))
    return ctx.finish(v, out)


def intt(root, inputs, ctx=None):
    """intt(root, input) src/fft/ntt.rs:51-68."""
    ctx = ctx or default_context()
    v = Vec(inputs)
    out, optr = ctx.out_like(v, _next_pow2(v.n) if v.n else 1)
    ctx.check(ctx.lib.zkb_intt(ctx.h, le16(root), v.ptr, v.n, optr))
    return ctx.finish(v, out)


def ntt_batch(root, columns, inverse=False, ctx=None):
    """`columns`: (batch, n, 2) array/tensor of independent columns of one length."""
    ctx = ctx or default_context()
    b, n = columns.shape[0], columns.shape[1]
    flat = Vec(columns.reshape(b * n, 2))
    n_out = _next_pow2(n)
    out, optr = ctx.out_like(flat, b * n_out)
    ctx.check(ctx.lib.zkb_ntt_batch(ctx.h, le16(root), 1 if inverse else 0, flat.ptr, n, n, optr, n_out, b))
    return out.reshape(b, n_out, 2)


def scale(coefficients, factor, ctx=None):
    """Polynomial::scale src/field/polynomial.rs:109-121."""
    ctx = ctx or default_context()
    v = Vec(coefficients)
    out, optr = ctx.out_like(v, v.n)
    ctx.check(ctx.lib.zkb_poly_scale(ctx.h, le16(factor), v.ptr, v.n, optr))
    return ctx.finish(v, out)


def fast_coset_evaluate(generator, root_order, offset, polynomial, ctx=None):
    """fast_coset_evaluate src/fft/ntt_arithmetics.rs:161-170 (the LDE): `polynomial` is the
    coefficient vector; returns root_order evaluations on offset*<generator>."""
    ctx = ctx or default_context()
    v = Vec(polynomial)
    out, optr = ctx.out_like(v, root_order)
    ctx.check(ctx.lib.zkb_coset_lde(ctx.h, le16(generator), root_order, le16(offset), v.ptr if v.n else None, v.n, optr))
    return ctx.finish(v, out)


def coset_lde_batch(generator, root_order, offset, columns, ctx=None):
    """LDE of `batch` coefficient columns ((batch, n, 2)) -> (batch, root_order, 2)."""
    ctx = ctx or default_context()
    b, n = columns.shape[0], columns.shape[1]
    flat = Vec(columns.reshape(b * n, 2))
    out, optr = ctx.out_like(flat, b * root_order)
    ctx.check(ctx.lib.zkb_coset_lde_batch(ctx.h, le16(generator), root_order, le16(offset), flat.ptr, n, n, optr, root_order, b))
    return out.reshape(b, root_order, 2)


def _binop(fn, ctx, head, lhs, rhs):
    l, r = Vec(lhs), Vec(rhs)
    assert l.kind != "cuda" and r.kind != "cuda", "fast_multiply / fast_coset_divide take host polynomials"
    out = np.empty((max(l.n + r.n, 1), 2), dtype=np.uint64)
    n_out = ctypes.c_size_t(0)
    ctx.check(fn(ctx.h, *head, l.ptr if l.n else None, l.n, r.ptr if r.n else None, r.n, out.ctypes.data, ctypes.byref(n_out)))
    res = out[:n_out.value]
    return unpack(res) if l.kind == "list" else res


def fast_multiply(root, root_order, lhs, rhs, ctx=None):
    """fast_multiply src/fft/ntt_arithmetics.rs:5-64."""
    ctx = ctx or default_context()
    return _binop(ctx.lib.zkb_poly_mul, ctx, (le16(root), root_order), lhs, rhs)


def fast_coset_divide(root, root_order, offset, lhs, rhs, ctx=None):
    """fast_coset_divide src/fft/ntt_arithmetics.rs:239-310."""
    ctx = ctx or default_context()
    return _binop(ctx.lib.zkb_coset_div, ctx, (le16(root), root_order, le16(offset)), lhs, rhs)
