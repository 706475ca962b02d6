"""The oracle's restatement of the hot path's CALLERS - MPolynomial, the matrix helpers,
Rescue-Prime, Stark::prove / verify and RPSSS - against every known answer the reference holds
for them (src/m_polynomial.rs:330-560, src/utils/matrix.rs:110-183,
src/rescue_prime/rescue_prime.rs:298-423, src/rpsss.rs:89-136).  CPU only."""
import json
import os

from oracle import field as F, poly as PL, proof_stream as PS
from oracle.mpoly import MPolynomial, inverse, poly_pow, rref, transpose
from oracle.rescue_prime import RescuePrime
from oracle.stark import RPSSS, Stark, deterministic_rng

P = F.P
KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def test_matrix_kats():
    assert transpose([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]) == [[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12]]
    m = [[1, 2, P - 1, P - 4], [2, 3, P - 1, P - 11], [P - 2, 0, P - 3, 22]]
    rref(m)
    assert m == [[1, 0, 0, P - 8], [0, 1, 0, 1], [0, 0, 1, P - 2]]
    assert inverse([[1, 2, 3], [0, 1, 4], [5, 6, 0]]) == [[P - 24, 18, 5], [20, P - 15, P - 4], [P - 5, 4, 1]]
    k = KATS["rescue_prime"]
    assert inverse([[int(x) for x in r] for r in k["MDS"]]) == [[int(x) for x in r] for r in k["MDS_inv"]]


def test_mpolynomial_kats():
    a = MPolynomial({(0, 1, 5): 17, (42, 1, 5): 5})
    b = MPolynomial({(42, 0): 8, (0, 0): P - 7})
    assert a * b == MPolynomial({(42, 1, 5): (136 + 5 * (P - 7)) % P, (0, 1, 5): 17 * (P - 7) % P, (84, 1, 5): 40})
    a = MPolynomial({(0, 1, 5): 17, (5, 23, 0): 5})
    b = MPolynomial({(42, 0): 8, (5, 23): 12})
    assert a + b == MPolynomial({(0, 1, 5): 17, (5, 23, 0): 17, (42, 0, 0): 8})
    assert -a == MPolynomial({(0, 1, 5): P - 17, (5, 23, 0): P - 5})
    assert a - b == MPolynomial({(0, 1, 5): 17, (5, 23, 0): (5 - 12) % P, (42, 0, 0): P - 8})
    assert MPolynomial.variables(3) == [MPolynomial({(1, 0, 0): 1}), MPolynomial({(0, 1, 0): 1}), MPolynomial({(0, 0, 1): 1})]
    assert MPolynomial.constant(0).is_zero() and not MPolynomial.constant(1).is_zero()
    m = MPolynomial({(1, 2, 5): 3, (5, 3, 4): 4})
    assert m ** 3 == MPolynomial({(11, 8, 13): 144, (3, 6, 15): 27, (7, 7, 14): 108, (15, 9, 12): 64})
    # lift / evaluate (m_polynomial.rs:437-470)
    # interpolate through (0,2), (1,5), (2,5): 2 + 4.5x - 1.5x^2
    half = F.inv(2)
    up = [2, 9 * half % P, (-3 * half) % P]
    assert [PL.evaluate(up, x) for x in (0, 1, 2)] == [2, 5, 5]
    assert PL.evaluate(up, 5) == MPolynomial.lift(up, 3).evaluate([0, 0, 0, 5])
    v = MPolynomial.variables(4)
    m1 = MPolynomial.constant(1) * v[0] + MPolynomial.constant(2) * v[1] + MPolynomial.constant(5) * (v[2] ** 3)
    m2 = MPolynomial.constant(1) * v[0] * v[3] + MPolynomial.constant(5) * (v[3] ** 3) + MPolynomial.constant(5)
    pt = [0, 5, 5, 2]
    assert m1.evaluate(pt) * m2.evaluate(pt) % P == (m1 * m2).evaluate(pt)
    assert (m1.evaluate(pt) + m2.evaluate(pt)) % P == (m1 + m2).evaluate(pt)
    # evaluate_symbolic (m_polynomial.rs:472-514), literal and grouped
    mp = MPolynomial({(0, 1, 5): 17, (6, 2, 13): 8})
    polys = [[5, 0, 2], [2, 6, 34], [8, 9, 10]]
    want = PL.add(PL.mul(PL.mul(PL.mul([17], poly_pow(polys[0], 0)), poly_pow(polys[1], 1)), poly_pow(polys[2], 5)),
                  PL.mul(PL.mul(PL.mul([8], poly_pow(polys[0], 6)), poly_pow(polys[1], 2)), poly_pow(polys[2], 13)))
    assert mp.evaluate_symbolic(polys) == want
    assert mp.evaluate_symbolic_grouped(polys) == want


def test_rescue_prime_kats():
    k = KATS["rescue_prime"]
    rp = RescuePrime(2, 1, 128, 27)
    assert rp.alpha == int(k["alpha"]) and rp.alpha_inv == int(k["alpha_inv"])
    assert rp.MDS == [[int(x) for x in r] for r in k["MDS"]]
    assert rp.MDS_inv == [[int(x) for x in r] for r in k["MDS_inv"]]
    assert rp.round_constants == [int(x) for x in k["round_constants"]]
    for x, h in k["hash"]:
        assert rp.hash(int(x)) == int(h)
    # rescue_prime.rs:333-423: the trace satisfies boundary + transition constraints; a disturbed one does not
    a, b = int(k["hash"][1][0]), int(k["hash"][1][1])
    trace = rp.trace(a)
    assert trace[0][0] == a and trace[-1][0] == b
    omicron = F.primitive_nth_root(1 << 119)
    tcs = rp.transition_constraints(omicron, 1 << 119)

    def ok(tr):
        for cycle, element, value in rp.boundary_constraints(b):
            if tr[cycle][element] != value:
                return False
        for i in range(len(tr) - 1):
            pt = [F.fpow(omicron, i)] + tr[i] + tr[i + 1]
            if any(tc.evaluate(pt) != 0 for tc in tcs):
                return False
        return True
    assert ok(trace)
    d = k["invalid_trace_delta"]
    trace[d["cycle"]][d["register"]] = (trace[d["cycle"]][d["register"]] + int(d["value"])) % P
    assert not ok(trace)


def test_evaluate_symbolic_grouped_equals_literal_on_the_air():
    """The Stark oracle evaluates the AIR symbolically with the grouped (fast) order of operations;
    on a short trace it must give the literal loop's polynomial."""
    rp = RescuePrime(2, 1, 128, 27)
    omicron = F.primitive_nth_root(64)
    tcs = rp.transition_constraints(omicron, 64)
    tp = [[(3 * i + 1) % P for i in range(5)], [(7 * i + 2) % P for i in range(5)]]
    point = [[0, 1]] + tp + [[c * F.fpow(omicron, i) % P for i, c in enumerate(t)] for t in tp]
    for tc in tcs:
        lit = tc.evaluate_symbolic(point)
        fast = tc.evaluate_symbolic_grouped(point)
        d = PL.degree(lit)
        assert d == PL.degree(fast) and lit[:d + 1] == fast[:d + 1]


def test_rpsss_sign_verify_round_trip():
    """src/rpsss.rs:103-135 at the tutorial parameters (4, 64, 128, 3): the signature verifies, a
    different document does not, and the signature is exactly 1,156,888 bytes (rpsss.rs:89)."""
    r = RPSSS(4, 64, 128, 3)
    assert (r.stark.omicron_domain_length, r.stark.fri_domain_length) == (1024, 4096)
    rng = deterministic_rng(b"rpsss-test")
    sk, pk = r.keygen(rng)
    sig = r.sign(sk, b"Hello, World!", rng)
    assert len(sig) == KATS["rpsss"]["signature_bytes"]
    assert r.verify(pk, b"Hello, World!", sig) is None
    assert r.verify(pk, b"Malicious document", sig) is not None
    assert r.verify((pk + 1) % P, b"Hello, World!", sig) is not None
    # deterministic given the byte source
    rng2 = deterministic_rng(b"rpsss-test")
    sk2, _ = r.keygen(rng2)
    assert sk2 == sk and r.sign(sk, b"Hello, World!", rng2) == sig


def test_stark_prove_verify_independent_stream():
    """stark.rs:810-880: Rescue-Prime hash-trace proof over an IndependentProofStream (configs[0])."""
    rp = RescuePrime(2, 1, 128, 27)
    stark = Stark(4, 64, 128, rp.m, rp.N + 1, 3)
    x = 0x1234567890ABCDEF
    out = rp.hash(x)
    tcs = rp.transition_constraints(stark.omicron, stark.omicron_domain_length)
    proof = stark.prove(rp.trace(x), tcs, rp.boundary_constraints(out), PS.IndependentProofStream(), deterministic_rng(b"s"))
    assert stark.verify(tcs, rp.boundary_constraints(out), PS.IndependentProofStream(PS.parse(proof))) is None
    assert stark.verify(tcs, rp.boundary_constraints((out + 1) % P), PS.IndependentProofStream(PS.parse(proof))) is not None
