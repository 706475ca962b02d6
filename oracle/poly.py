"""Univariate polynomial helpers and the divide-and-conquer domain algorithms (oracle; test
infrastructure only).  Restates src/field/polynomial.rs (degree :46-63, evaluate :75-96,
divide_with_rem :179-224, + :254-281, - :283-288, * :290-314, % :316-326) and
src/fft/ntt_arithmetics.rs fast_zerofier :66-108, fast_evaluate_domain :110-159,
fast_interpolate_domain :172-237.  Coefficient vectors are kept exactly as the reference keeps
them (trailing zeros are NOT trimmed), because callers compare / serialise them as they are."""
from . import field as F
from . import ntt as N

P = F.P


def degree(p):
    return N.degree(p)


def add(a, b):
    if degree(a) is None:                      # polynomial.rs:256-261: a zero operand returns the other AS IS
        return list(b)
    if degree(b) is None:
        return list(a)
    out = [0] * max(len(a), len(b))
    for i, c in enumerate(a):
        out[i] = (out[i] + c) % P
    for i, c in enumerate(b):
        out[i] = (out[i] + c) % P
    return out


def neg(a):
    return [(-c) % P for c in a]


def sub(a, b):
    return add(a, neg(b))


def mul(a, b):
    if len(a) == 0 or len(b) == 0:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % P
    return out


def leading(p):
    d = degree(p)
    return p[d] if d is not None else (p[-1] if p else None)


def divide_with_rem(num, den):
    dd = degree(den)
    assert dd is not None, "Denominator is zero or empty"
    nd = degree(num)
    if nd is None or nd < dd:
        return [], list(num)
    rem = list(num)
    steps = nd - dd + 1
    quo = [0] * steps
    lead = leading(den)
    for _ in range(steps):
        rd = degree(rem)
        if rd is None or rd < dd:
            break
        coef = F.div(leading(rem), lead)
        shift = rd - dd
        rem = sub(rem, mul([0] * shift + [coef], den))
        quo[shift] = coef
    return quo, rem


def rem(a, b):
    return divide_with_rem(a, b)[1]


def evaluate(p, x):
    value, xi = 0, 1
    for c in p:
        value = (value + c * xi) % P
        xi = xi * x % P
    return value


def zerofier_domain(domain):
    acc = [1]
    for d in domain:
        acc = mul(acc, [(-d) % P, 1])
    return acc


def fast_zerofier(root, root_order, domain):
    N._check_root(root, root_order)

    def inner(dom):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [(-dom[0]) % P, 1]
        half = len(dom) // 2
        return N.fast_multiply(root, root_order, inner(dom[:half]), inner(dom[half:]))
    return inner(list(domain))


def fast_evaluate_domain(root, root_order, poly, domain):
    N._check_root(root, root_order)

    def inner(p, dom):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [evaluate(p, dom[0])]
        half = len(dom) // 2
        left = fast_zerofier(root, root_order, dom[:half])
        right = fast_zerofier(root, root_order, dom[half:])
        return inner(rem(p, left), dom[:half]) + inner(rem(p, right), dom[half:])
    return inner(list(poly), list(domain))


def fast_interpolate_domain(root, root_order, domain, values):
    N._check_root(root, root_order)
    assert len(domain) == len(values)

    def inner(dom, vals):
        if len(dom) == 0:
            return []
        if len(dom) == 1:
            return [vals[0]]
        half = len(dom) // 2
        lz = fast_zerofier(root, root_order, dom[:half])
        rz = fast_zerofier(root, root_order, dom[half:])
        lo = fast_evaluate_domain(root, root_order, rz, dom[:half])
        ro = fast_evaluate_domain(root, root_order, lz, dom[half:])
        lt = [F.div(vals[i], d) for i, d in enumerate(lo)]
        rt = [F.div(vals[i + half], d) for i, d in enumerate(ro)]
        li = inner(dom[:half], lt)
        ri = inner(dom[half:], rt)
        return add(mul(li, rz), mul(ri, lz))
    return inner(list(domain), list(values))
