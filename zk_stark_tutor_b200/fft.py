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


# ---- divide-and-conquer algorithms over arbitrary domains (src/fft/ntt_arithmetics.rs:66-237) ----
# In the reference these recurse over fast_multiply (NTT: on the GPU here) and the scalar
# Polynomial helpers `%`, `*`, `+`, `evaluate`, `/` (src/field/polynomial.rs), which stay scalar
# host code exactly as in the reference; coefficient vectors keep the reference's trailing zeros.
_P = 1 + 407 * (1 << 119)


def _degree(p):
    d = None
    for i, c in enumerate(p):
        if c != 0:
            d = i
    return d


def _padd(a, b):                                   # polynomial.rs:254-281
    if _degree(a) is None:
        return list(b)
    if _degree(b) is None:
        return list(a)
    out = [0] * max(len(a), len(b))
    for i, c in enumerate(a):
        out[i] = (out[i] + c) % _P
    for i, c in enumerate(b):
        out[i] = (out[i] + c) % _P
    return out


def _pmul(a, b):                                   # polynomial.rs:290-314 (schoolbook)
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % _P
    return out


def _prem(num, den):                               # polynomial.rs:179-224, :316-326
    dd = _degree(den)
    assert dd is not None, "Denominator is zero or empty"
    nd = _degree(num)
    if nd is None or nd < dd:
        return list(num)
    rem = list(num)
    lead_inv = pow(den[dd], _P - 2, _P)
    for _ in range(nd - dd + 1):
        rd = _degree(rem)
        if rd is None or rd < dd:
            break
        coef = rem[rd] * lead_inv % _P
        shift = rd - dd
        sub = [0] * shift + [coef * c % _P for c in den]          # (coef * x^shift) * den
        rem = _padd(rem, [(-c) % _P for c in sub])
    return rem


def _peval(p, x):                                  # polynomial.rs:75-96
    value, xi = 0, 1
    for c in p:
        value = (value + c * xi) % _P
        xi = xi * x % _P
    return value


def fast_zerofier(root, root_order, domain, ctx=None):
    """fast_zerofier src/fft/ntt_arithmetics.rs:66-108."""
    ctx = ctx or default_context()
    from .context import ZkbError
    if pow(root, root_order, _P) != 1 or pow(root, root_order // 2, _P) == 1:
        raise ZkbError(-6, "supplied root %d is not a primitive root of root_order %d" % (root, root_order))

    def inner(dom):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [(-dom[0]) % _P, 1]
        half = len(dom) // 2
        return fast_multiply(root, root_order, inner(dom[:half]), inner(dom[half:]), ctx)
    return inner(list(domain))


def fast_evaluate_domain(root, root_order, polynomial, domain, ctx=None):
    """fast_evaluate_domain src/fft/ntt_arithmetics.rs:110-159."""
    ctx = ctx or default_context()

    def inner(p, dom):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [_peval(p, dom[0])]
        half = len(dom) // 2
        left = fast_zerofier(root, root_order, dom[:half], ctx)
        right = fast_zerofier(root, root_order, dom[half:], ctx)
        return inner(_prem(p, left), dom[:half]) + inner(_prem(p, right), dom[half:])
    fast_zerofier(root, root_order, [], ctx)       # the root-order asserts (:116-129)
    return inner(list(polynomial), list(domain))


def fast_interpolate_domain(root, root_order, domain, values, ctx=None):
    """fast_interpolate_domain src/fft/ntt_arithmetics.rs:172-237."""
    ctx = ctx or default_context()
    assert len(domain) == len(values)

    def inner(dom, vals):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [vals[0]]
        half = len(dom) // 2
        lz = fast_zerofier(root, root_order, dom[:half], ctx)
        rz = fast_zerofier(root, root_order, dom[half:], ctx)
        lo = fast_evaluate_domain(root, root_order, rz, dom[:half], ctx)
        ro = fast_evaluate_domain(root, root_order, lz, dom[half:], ctx)
        lt = [vals[i] * pow(d, _P - 2, _P) % _P for i, d in enumerate(lo)]
        rt = [vals[i + half] * pow(d, _P - 2, _P) % _P for i, d in enumerate(ro)]
        li = inner(dom[:half], lt)
        ri = inner(dom[half:], rt)
        return _padd(_pmul(li, rz), _pmul(ri, lz))
    fast_zerofier(root, root_order, [], ctx)
    return inner(list(domain), list(values))
