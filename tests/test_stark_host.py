"""Host-side algorithms of zk_stark_tutor_b200/stark.py (the device Stark prover) on the CPU, with the oracle's
NTT products injected where the product code calls the GPU: the subgroup-prefix interpolation that replaces the
reference's divide-and-conquer fast_interpolate_domain for trace columns (stark.rs:305-326), the small Lagrange
interpolation of boundary conditions, and the degree bookkeeping (stark.rs:115-258) against the oracle's."""
import random

import pytest

from oracle import field as F, ntt as N, poly as PL
from oracle.stark import RPSSS
from zk_stark_tutor_b200 import stark as S

P = F.P


@pytest.mark.parametrize("L,Nn", [(284, 1024), (1, 4), (2, 4), (5, 8), (7, 8), (8, 8), (100, 128), (129, 256)])
def test_prefix_interpolator_equals_fast_interpolate_domain(L, Nn):
    rnd = random.Random(L * 1000 + Nn)
    w, big = F.primitive_nth_root(Nn), F.primitive_nth_root(4 * Nn)
    dom = [F.fpow(w, i) for i in range(L)]
    Z = PL.fast_zerofier(w, Nn, dom) if 1 < L < Nn else [(-1) % P, 1]
    it = S.PrefixInterpolator(L, Nn, Z, mul=lambda a, b: N.fast_multiply(big, 4 * Nn, a, b), intt=lambda v: N.intt(w, v))
    for vals in ([rnd.randrange(P) for _ in range(L)], [0] * L, [7] * L, [P - 1] + [0] * (L - 1)):
        got = it(vals)
        assert len(got) <= max(L, Nn if L == Nn else L)
        assert [PL.evaluate(got, x) for x in dom] == [v % P for v in vals]
        assert (PL.degree(got) or 0) < L
        if L > 1:
            want = PL.fast_interpolate_domain(w, Nn, dom, vals)
            d = PL.degree(want)
            assert PL.degree(got) == d and (d is None or got[:d + 1] == want[:d + 1])


def test_prefix_interpolator_array_mode():
    """the GPU path keeps coefficient vectors as (n, 2) uint64 arrays between library calls"""
    from oracle import cbind as C
    L, Nn = 284, 1024
    rnd = random.Random(11)
    w, big = F.primitive_nth_root(Nn), F.primitive_nth_root(4 * Nn)
    dom = [F.fpow(w, i) for i in range(L)]
    mul = lambda a, b: N.fast_multiply(big, 4 * Nn, a, b)                                     # noqa: E731
    amul = lambda a, b: C.to_arr(N.fast_multiply(big, 4 * Nn, C.from_arr(a), C.from_arr(b)))  # noqa: E731
    aintt = lambda v: C.to_arr(N.intt(w, C.from_arr(v)))                                      # noqa: E731
    it = S.PrefixInterpolator(L, Nn, PL.fast_zerofier(w, Nn, dom), mul, lambda v: N.intt(w, v))
    vals = [rnd.randrange(P) for _ in range(L)]
    want = it(vals)
    assert it.use_arrays(amul, aintt)(vals) == want
    assert [PL.evaluate(want, x) for x in dom[:7]] == vals[:7]


def test_sample_is_field_sample():
    for data in (bytes.fromhex("6c9c4992"), bytes.fromhex("ac4cd3be"), bytes(range(17)), b"\xff" * 40, b""):
        assert S.Stark.sample(data) == F.sample(data)


def test_lagrange_interpolate_small_domains():
    rnd = random.Random(3)
    w = F.primitive_nth_root(1024)
    for k in (1, 2, 3, 8):
        dom = [F.fpow(w, rnd.randrange(1024)) for _ in range(k)]
        if len(set(dom)) < k:
            continue
        vals = [rnd.randrange(P) for _ in range(k)]
        got = S.lagrange_interpolate(dom, vals)
        assert [PL.evaluate(got, x) for x in dom] == vals
        if k > 1:
            want = PL.fast_interpolate_domain(w, 1024, dom, vals)
            d = PL.degree(want)
            assert PL.degree(got) == d and got[:d + 1] == want[:d + 1]


def test_degree_bookkeeping_equals_oracle():
    """stark.rs:115-258 restated in the product prover == the oracle's restatement, on the Rescue-Prime AIR"""
    r = RPSSS(4, 64, 128, 3)
    tcs = r.transition_constraints()
    dev = S.Stark.__new__(S.Stark)                       # the bookkeeping needs no device
    dev.original_trace_length, dev.num_randomizers, dev.num_registers = r.stark.original_trace_length, r.stark.num_randomizers, r.stark.num_registers
    assert dev.transition_degree_bounds(tcs) == r.stark.transition_degree_bounds(tcs)
    assert dev.transition_quotient_degree_bounds([t.dictionary for t in tcs]) == r.stark.transition_quotient_degree_bounds(tcs)
    assert dev.max_degree(tcs) == r.stark.max_degree(tcs) == 1023
    assert S._bit_count(0) == 1 and S._bit_count(1) == 1 and S._bit_count(852) == 10


def test_sample_many_and_zerofier_of(monkeypatch):
    import os
    from zk_stark_tutor_b200.context import unpack
    buf = bytes([0]) + P.to_bytes(16, "big") + bytes([9]) + (P + 5).to_bytes(16, "big") + bytes([0]) + ((1 << 128) - 1).to_bytes(16, "big") \
        + bytes([1]) + (P - 1).to_bytes(16, "big") + os.urandom(17 * 500)
    monkeypatch.setattr(os, "urandom", lambda k: buf[:k])
    want = [S.Stark.sample(buf[17 * i:17 * i + 17]) for i in range(504)]
    assert want[:4] == [0, 5, (1 << 128) - 1 - P, P - 1]
    assert unpack(S.sample_many(os.urandom, 504)) == want                     # one bulk draw, vectorised reduction
    calls = iter(buf[17 * i:17 * i + 17] for i in range(504))
    assert unpack(S.sample_many(lambda k: next(calls), 504)) == want          # any other byte source: element by element
    w = F.primitive_nth_root(1024)
    pts = [F.fpow(w, i) for i in range(27)]
    assert S.zerofier_of(pts) == PL.fast_zerofier(w, 1024, pts)
    assert S.zerofier_of([]) == [1]


def test_flatten_constraints_matches_mpolynomial_evaluate():
    """air.flatten_constraints: the flat (counts, coefs, exps) arrays handed to zkb_air_create evaluate to MPolynomial::evaluate
    (m_polynomial.rs:97-126) at random points - short keys are zero padded, zero-coefficient entries are kept."""
    from zk_stark_tutor_b200.air import flatten_constraints
    from zk_stark_tutor_b200.context import unpack
    from oracle.mpoly import MPolynomial
    rnd = random.Random(21)
    nr = 2
    tcs = [MPolynomial({(0,): 7, (3, 1): 5, (0, 0, 2, 0, 1): P - 1, (27, 0, 0, 3): 0}), MPolynomial({(1, 1, 1, 1, 1): 9}), MPolynomial({})]
    counts, coefs, exps = flatten_constraints(tcs, nr)
    assert list(counts) == [4, 1, 0] and exps.shape == (5, 5) and coefs.shape == (5, 2)
    cf = unpack(coefs)
    for _ in range(5):
        pt = [rnd.randrange(P) for _ in range(1 + 2 * nr)]
        t = 0
        for j, tc in enumerate(tcs):
            acc = 0
            for k in range(int(counts[j])):
                term = cf[t]
                for v, e in zip(pt, exps[t]):
                    term = term * pow(v, int(e), P) % P
                acc = (acc + term) % P
                t += 1
            assert acc == tc.evaluate(pt)
    with pytest.raises(ValueError):
        flatten_constraints([{(0, 0, 0, 0, 0, 1): 3}], nr)          # a sixth variable with 2 registers
    assert flatten_constraints([{(0, 0, 0, 0, 0, 0): 3}], nr)[2].shape == (1, 5)   # trailing zero exponents beyond the variables are fine
