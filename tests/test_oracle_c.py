"""The C oracle (fast zo_* and faithful zr_*) against the Python oracle.  CPU only."""
import random

import numpy as np

from oracle import cbind as C
from oracle import field as F
from oracle import merkle as M
from oracle import ntt as N
from oracle.fri import FRI

P = F.P
rnd = random.Random(1234)


def rvals(n):
    return [rnd.randrange(P) for _ in range(n)]


def test_scalar_ops():
    edge = [0, 1, 2, P - 1, P - 2, (1 << 127), (1 << 64) - 1, 1 << 64, F.GENERATOR]
    vals = edge + rvals(50)
    for a in vals:
        for b in vals[:12]:
            assert C.scalar("zo_mul", a, b) == a * b % P
            assert C.scalar("zr_mul", a, b) == a * b % P
    for a in vals:
        if a:
            assert C.scalar("zo_inv", a) == F.inv(a)
            assert C.scalar("zr_inv", a) == F.inv(a)
    assert C.scalar("zo_pow", 6534789852937546098, 501209126122) == 256557788041265930815463337858691703671
    assert C.scalar("zr_pow", 6534789852937546098, 501209126122) == 256557788041265930815463337858691703671
    assert C.scalar("zr_pow", 5, 0) == 1


def test_synth_matches_python():
    assert C.from_arr(C.synth(0x5EED0002, 300, start=5)) == F.synth_elements(0x5EED0002, 300, start=5)


def test_ntt_vs_python():
    for n_in in (1, 2, 3, 5, 16, 100, 1024):
        xs = rvals(n_in)
        n = N.next_pow2(n_in)
        w = F.primitive_nth_root(max(n, 2))
        want = N.ntt(w, xs)
        assert C.from_arr(C.ntt(w, C.to_arr(xs))) == want
        assert C.from_arr(C.ntt(w, C.to_arr(xs), faithful=True)) == want
        assert C.from_arr(C.ntt(w, C.to_arr(want), inverse=True)) == N.intt(w, want)


def test_lde_merkle_fold_vs_python():
    n, ef = 256, 4
    w = F.primitive_nth_root(n)
    coeffs = rvals(n // ef - 3)
    want = N.fast_coset_evaluate(w, n, F.GENERATOR, coeffs)
    got = C.coset_lde(w, n, F.GENERATOR, C.to_arr(coeffs))
    assert C.from_arr(got) == want
    assert C.from_arr(C.coset_lde(w, n, F.GENERATOR, C.to_arr(coeffs), faithful=True)) == want
    root, nodes = C.merkle(got, want_nodes=True)
    assert root == M.commit(want)
    assert C.merkle(got, faithful=True) == root
    lv = M.tree_levels(want)
    flat = b"".join(b"".join(level) for level in lv)
    assert nodes.tobytes() == flat
    alpha = rvals(1)[0]
    wantf = FRI.fold(want, alpha, F.GENERATOR, w)
    assert C.from_arr(C.fri_fold(got, alpha, F.GENERATOR, w)) == wantf
    assert C.from_arr(C.fri_fold(got, alpha, F.GENERATOR, w, faithful=True)) == wantf
    # small/edge values through the leaf encoder
    edge = [0, 1, 9, 10, 11, 5462, 10**18, 10**19 - 1, 10**19, 10**38, P - 1, (1 << 64), (1 << 64) - 1] + [3] * 3
    assert C.merkle(C.to_arr(edge)) == M.commit(edge)


def test_big_fold_chunking():
    n = 1 << 15
    w = F.primitive_nth_root(n)
    cw = C.synth(7, n)
    alpha = 0x1234567890ABCDEF1234567890ABCDEF % P
    got = C.from_arr(C.fri_fold(cw, alpha, F.GENERATOR, w))
    vals = C.from_arr(cw)
    half = n // 2
    for i in (0, 1, 4095, 4096, 4097, half - 1):
        ax = F.div(alpha, F.mul(F.GENERATOR, F.fpow(w, i)))
        want = F.mul(F.inv(2), F.add(F.mul(F.add(1, ax), vals[i]), F.mul(F.sub(1, ax), vals[half + i])))
        assert got[i] == want


def test_fastfri_matches_literal_restatement():
    """oracle.fastfri (C kernels, trees built once) == oracle.fri.FRI.prove (literal fri.rs)."""
    from oracle import fastfri, proof_stream as PS
    for n, ncc, doc in ((256, 17, None), (2048, 64, None), (1024, 20, b"doc")):
        w = F.primitive_nth_root(n)
        coeffs = rvals(n // 4)
        cw = N.fast_coset_evaluate(w, n, F.GENERATOR, coeffs)
        fri = FRI(F.GENERATOR, w, n, 4, ncc)
        mk = (lambda: PS.SignatureProofStream(doc)) if doc else PS.IndependentProofStream
        ps1, ps2 = mk(), mk()
        top1 = fri.prove(cw, ps1)
        top2, _, _ = fastfri.prove(fri, C.to_arr(cw), ps2)
        assert top1 == top2
        assert ps1.digest() == ps2.digest()
        assert fri.verify(ps2, []) is None


def test_bench_digest_file_is_what_the_oracle_computes():
    """tests/golden/bench_digests.json (bench.py's pre-timing parity checks and the full-size GPU tests rely on it) against the oracle run
    here, at the sizes that take seconds: configs[2] at 2^16 / 2^18 and the configs[4] checksums at 2^20."""
    import hashlib
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    sys.path.insert(0, os.path.join(root, "tools"))
    import make_bench_digests as M
    from golden_ntt import check_against_golden_ntt
    gold = json.load(open(os.path.join(root, "tests", "golden", "bench_digests.json")))
    for log_n in (16, 18):
        assert M.commit_digest(log_n, M.SEED) == gold["configs2"][str(log_n)]
    n = 1 << 20
    check_against_golden_ntt(C.ntt(F.primitive_nth_root(n), C.synth(M.SEED_NTT, n)), 20)
    assert M.checksums(C.synth(1, 1000), first_index=5, step=3)["sum_lo"] == int(C.synth(1, 1000)[:, 0].sum(dtype="uint64"))
