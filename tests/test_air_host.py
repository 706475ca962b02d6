"""csrc/air.cuh on the CPU: the evaluation-form transition quotients + nonlinear combination
(zkb_air_combination's point body and term grouping, compiled for the host by tests/cpp/air_host.cpp)
against the oracle's coefficient-form Stark::prove (stark.rs:388-519) on the real Rescue-Prime AIR at the
tutorial's signature parameters: the combined codeword and both transition-quotient codewords must be
identical value for value.  The device build of the same body is checked by tests/test_gpu_air.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from air_common import flatten, prover_intermediates
from oracle import cbind as C, field as F, ntt as N, poly as PL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("air") / "libair_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "cpp", "air_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


def test_air_point_body_equals_coefficient_form_prover(lib):
    m = prover_intermediates()
    st, nr, nc, n = m["stark"], m["nr"], m["nc"], m["n"]
    counts, coefs, exps = flatten(m["tcs"], nr)
    lde = lambda p: C.coset_lde(st.omega, n, st.generator, C.to_arr(list(p)))        # noqa: E731
    bq = np.concatenate([C.to_arr(cw) for cw in m["bq_cws"]])
    zb = np.concatenate([lde(z) for z in m["zerofiers"]])
    ib = np.concatenate([lde(p) for p in m["interpolants"]])
    tz = lde(m["tz"])
    out = np.zeros((n, 2), dtype=np.uint64)
    tq = np.zeros((nc * n, 2), dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                                   # noqa: E731
    rc = lib.air_host_combination(ctypes.c_uint64(n), ctypes.c_uint64(st.expansion_factor), nr, nc, 1, p(counts), p(coefs), p(exps),
                                  p(bq), p(C.to_arr(m["rnd_cw"])), p(zb), p(ib), p(tz), p(C.to_arr(m["weights"])),
                                  p(np.asarray(m["shifts"], dtype=np.uint32)), p(C.to_arr([st.generator])), p(C.to_arr([st.omega])), p(out), p(tq))
    assert rc > 0, rc
    # the grouping collapses the ~600 dictionary terms into a handful of univariate-in-x groups
    assert rc < sum(counts)
    assert C.from_arr(out) == m["combined"], "evaluation-form combination differs from LDE(weighted sum of terms)"
    for j in range(nc):
        want = N.fast_coset_evaluate(st.omega, n, st.generator, m["tq_polys"][j])
        assert C.from_arr(tq[j * n:(j + 1) * n]) == want, "transition quotient %d" % j


def test_air_point_flags_division_by_zero(lib):
    """If Z_T vanishes on the domain the reference's division panics (field_element.rs:85): the body reports it."""
    n, nr, nc = 8, 1, 1
    counts = np.asarray([1], dtype=np.uint32)
    coefs = C.to_arr([1])
    exps = np.asarray([[0, 1, 0]], dtype=np.uint32)
    ones = C.to_arr([1] * n)
    tz = C.to_arr([1, 2, 0, 4, 5, 6, 7, 8])
    out = np.zeros((n, 2), dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                                   # noqa: E731
    rc = lib.air_host_combination(ctypes.c_uint64(n), ctypes.c_uint64(2), nr, nc, 1, p(counts), p(coefs), p(exps), p(ones), p(ones), p(ones), p(ones),
                                  p(tz), p(C.to_arr([1, 1, 1, 1, 1])), p(np.asarray([0, 0], dtype=np.uint32)),
                                  p(C.to_arr([F.GENERATOR])), p(C.to_arr([F.primitive_nth_root(n)])), p(out), None)
    assert rc == -7


def test_boundary_quotient_body_and_batch_strides(lib):
    """k_boundary_quotient's body: (t - I) / Z_B on the coset == LDE of the coefficient-form boundary quotient
    (stark.rs:331-360); and two instances through the instance strides of the combination body."""
    m = prover_intermediates()
    st, nr, nc, n = m["stark"], m["nr"], m["nc"], m["n"]
    lde = lambda p: C.coset_lde(st.omega, n, st.generator, C.to_arr(list(p)))        # noqa: E731
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                                   # noqa: E731
    for s in range(nr):
        # t = bq * Z_B + I as polynomials -> its LDE is the trace codeword
        t_poly = PL.add(N.fast_multiply(st.omega, n, m["bq_polys"][s], m["zerofiers"][s]), m["interpolants"][s])
        out = np.zeros((n, 2), dtype=np.uint64)
        lib.air_host_boundary_quotient(ctypes.c_uint64(n), p(lde(t_poly)), p(lde(m["interpolants"][s])), p(lde(m["zerofiers"][s])), p(out))
        assert C.from_arr(out) == m["bq_cws"][s]
    # batch of 2: instance 1 = instance 0 with doubled weights -> combination doubles, quotients unchanged
    counts, coefs, exps = flatten(m["tcs"], nr)
    bq1 = np.concatenate([C.to_arr(cw) for cw in m["bq_cws"]])
    bq = np.concatenate([bq1, bq1])
    rnd = np.concatenate([C.to_arr(m["rnd_cw"])] * 2)
    zb = np.concatenate([lde(z) for z in m["zerofiers"]])
    ib1 = np.concatenate([lde(q) for q in m["interpolants"]])
    ib = np.concatenate([ib1, ib1])
    w2 = [2 * w % F.P for w in m["weights"]]
    out = np.zeros((2 * n, 2), dtype=np.uint64)
    rc = lib.air_host_combination(ctypes.c_uint64(n), ctypes.c_uint64(st.expansion_factor), nr, nc, 2, p(counts), p(coefs), p(exps),
                                  p(bq), p(rnd), p(zb), p(ib), p(lde(m["tz"])), p(C.to_arr(m["weights"] + w2)),
                                  p(np.asarray(m["shifts"], dtype=np.uint32)), p(C.to_arr([st.generator])), p(C.to_arr([st.omega])), p(out), None)
    assert rc > 0
    got = C.from_arr(out)
    assert got[:n] == m["combined"] and got[n:] == [2 * v % F.P for v in m["combined"]]


def test_prefix_tables_host_arithmetic(lib):
    """csrc/prefix.cuh: Z = prod (X - root^i) equals fast_zerofier's polynomial, and rev(Z) * g == 1 mod X^m"""
    for L, Nn in ((284, 1024), (5, 8), (2, 4), (100, 128)):
        w = F.primitive_nth_root(Nn)
        mm = Nn - L
        Z = np.zeros((L + 1, 2), dtype=np.uint64)
        g = np.zeros((max(mm, 1), 2), dtype=np.uint64)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                               # noqa: E731
        lib.prefix_host_tables(p(C.to_arr([w])), ctypes.c_uint64(L), ctypes.c_uint64(mm), p(Z), p(g))
        Zl = C.from_arr(Z)
        want = PL.fast_zerofier(w, Nn, [F.fpow(w, i) for i in range(L)])
        assert Zl == want[:L + 1] and not any(want[L + 1:])
        prod = PL.mul(Zl[::-1][:mm], C.from_arr(g)[:mm])[:mm]
        assert prod == [1] + [0] * (mm - 1)
