"""NTT / iNTT / coset LDE / NTT-based polynomial arithmetic (oracle; test infra only).

Restates src/utils/bit_reverse_copy.rs:3-34, src/fft/ntt.rs:7-68,
src/field/polynomial.rs:46-63 (degree) and :109-121 (scale),
src/fft/ntt_arithmetics.rs:5-64 (fast_multiply), :161-170 (fast_coset_evaluate),
:239-310 (fast_coset_divide).  Polynomials are plain lists of canonical ints,
lowest-degree coefficient first.
"""
from . import field as F


def next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def bit_reverse_copy(xs):
    # bit_reverse_copy.rs:3-34: <2 elements returned as is; else zero-pad to the next
    # power of two, then out[rev(k)] = in[k] over log2(n) bits.
    if len(xs) < 2:
        return list(xs)
    n = next_pow2(len(xs))
    bits = n.bit_length() - 1
    out = [0] * n
    for k, v in enumerate(xs):
        out[int(format(k, "0%db" % bits)[::-1], 2)] = v
    return out


def ntt(root, xs):
    # ntt.rs:7-49: iterative radix-2 DIT, natural-order output
    # out[k] = sum_i in[i] * root^(i*k).  Empty input panics in the reference (ntt.rs:11).
    assert len(xs) >= 1, "ntt: empty input (reference panics on inputs[0])"
    a = bit_reverse_copy(xs)
    n = len(a)
    pw = [1] * (n // 2)
    for k in range(1, n // 2):
        pw[k] = pw[k - 1] * root % F.P
    size = 1
    while size < n:
        size <<= 1
        half, step = size >> 1, n // size
        for i in range(0, n, size):
            k = 0
            for j in range(i, i + half):
                even, odd = a[j], a[j + half] * pw[k] % F.P
                a[j] = (even + odd) % F.P
                a[j + half] = (even - odd) % F.P
                k += step
    return a


def intt(root, xs):
    # ntt.rs:51-68: identity for len < 2; otherwise ntt(root^-1) scaled by
    # (next_pow2(len))^-1.
    if len(xs) < 2:
        return list(xs)
    ninv = F.inv(next_pow2(len(xs)) % F.P)
    return [ninv * v % F.P for v in ntt(F.inv(root), xs)]


def degree(poly):
    # polynomial.rs:46-63: index of the last non-zero coefficient, None if all zero / empty.
    d = None
    for i, c in enumerate(poly):
        if c != 0:
            d = i
    return d


def scale(poly, factor):
    # polynomial.rs:109-121: c_i <- factor^i * c_i
    out, pw = [], 1
    for c in poly:
        out.append(pw * c % F.P)
        pw = pw * factor % F.P
    return out


def fast_coset_evaluate(generator, root_order, offset, poly):
    # ntt_arithmetics.rs:161-170 (the LDE): scale(offset), zero-pad to root_order, ntt.
    assert len(poly) <= root_order, "usize underflow in the reference (coeffs.len() > root_order)"
    coeffs = scale(poly, offset)
    coeffs += [0] * (root_order - len(coeffs))
    return ntt(generator, coeffs)


def _check_root(root, root_order):
    assert F.fpow(root, root_order) == 1, "supplied root does not have supplied root_order"
    assert F.fpow(root, root_order // 2) != 1, "supplied root is not a primitive of root_order"


def fast_multiply(root, root_order, lhs, rhs):
    # ntt_arithmetics.rs:5-64
    _check_root(root, root_order)
    dl, dr = degree(lhs), degree(rhs)
    if dl is None or dr is None:
        return []
    deg = dl + dr
    order = root_order
    while deg < order // 2:
        root = root * root % F.P
        order //= 2
    # `extend(len..order)` is a no-op when the operand is already longer than `order`;
    # ntt() then pads to its own next power of two (reference behaviour, kept).
    pad = lambda p: list(p) + [0] * max(0, order - len(p))
    a, b = ntt(root, pad(lhs)), ntt(root, pad(rhs))
    had = [a[i] * b[i] % F.P for i in range(order)]
    return intt(root, had)[:deg + 1]


def fast_coset_divide(root, root_order, offset, lhs, rhs):
    # ntt_arithmetics.rs:239-310
    _check_root(root, root_order)
    dr = degree(rhs)
    assert dr is not None, "cannot divide by zero polynomial"
    dl = degree(lhs)
    if dl is None:
        return []
    assert dl >= dr, "cannot divide by polynomial of larger degree"
    deg = max(dl, dr)
    order = root_order
    while deg < order // 2:
        root = root * root % F.P
        order //= 2
    def ev(p):
        p = scale(p, offset)
        return ntt(root, p + [0] * max(0, order - len(p)))
    a, b = ev(lhs), ev(rhs)
    quo = [F.div(a[i], b[i]) for i in range(order)]
    coeffs = intt(root, quo)[:dl - dr + 1]
    return scale(coeffs, F.inv(offset))
