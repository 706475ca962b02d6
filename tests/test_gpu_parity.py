"""GPU parity: the CUDA path (through the C ABI of libzkb200.so and its Python mirror of the
reference's API) against the oracle, bit-exact.  Run on the B200 box: pytest -m gpu.

Layout follows the reference's own tests: known-answer vectors first
(src/fft/ntt.rs:78-130, src/merkle_root.rs:107-244), then randomised fast-vs-schoolbook
properties (src/fft/ntt_arithmetics.rs:356-517), then FRI round trips
(src/fri.rs:451-531), plus full-size cases from BASELINE.json's configs."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import zk_stark_tutor_b200 as zk
from zk_stark_tutor_b200 import proof_stream as ZPS
from oracle import cbind as C, field as F, merkle as M, ntt as N, proof_stream as PS, fastfri
from oracle.fri import FRI as OFRI

pytestmark = pytest.mark.gpu
P = F.P
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))
VEC = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
rnd = random.Random(99)


def rvals(n):
    return [rnd.randrange(P) for _ in range(n)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)            # raises if the CUDA library / device is missing: no fallback
    yield c
    c.close()


def cuda(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


# ---------------------------------------------------------------- NTT ----------------------
def test_ntt_reference_kats(ctx):
    w = F.primitive_nth_root(16)
    k = KATS["ntt16"]
    assert zk.ntt(w, [int(v) for v in k["in"]], ctx) == [int(v) for v in k["out"]]
    k = KATS["intt16"]
    assert zk.ntt(w, [int(v) for v in k["values"]], ctx) == [int(v) for v in k["coeffs"]]
    assert zk.intt(w, [int(v) for v in k["coeffs"]], ctx) == [int(v) for v in k["values"]]


@pytest.mark.parametrize("log_n", list(range(1, 15)) + [16, 17, 20, 21, 22, 23])
def test_ntt_vs_oracle(ctx, log_n):
    n = 1 << log_n
    x = C.synth(0x5EED0002, n)
    w = F.primitive_nth_root(n)
    want = C.ntt(w, x)
    got = zk.ntt(w, x, ctx)
    assert np.array_equal(got, want)
    back = zk.intt(w, got, ctx)
    assert np.array_equal(back, x)
    assert np.array_equal(zk.intt(w, x, ctx), C.ntt(w, x, inverse=True))
    for g in VEC["ntt"]:
        if g["log_n"] == log_n:
            assert sha(got) == g["sha256_out"]


def test_ntt_edge_values(ctx):
    n = 64
    w = F.primitive_nth_root(n)
    for xs in ([0] * n, [P - 1] * n, [1] + [0] * (n - 1), [P - 1, 1] * (n // 2), [(1 << 96) - 1] * n, [1 << 127] * n):
        assert zk.ntt(w, xs, ctx) == N.ntt(w, xs)


@pytest.mark.parametrize("n_in", [1, 2, 3, 5, 7, 100, 1000, 4097, 70000, (1 << 20) + 1])
def test_ntt_ragged_lengths_are_zero_padded(ctx, n_in):
    x = C.synth(11, n_in)
    n = C.next_pow2(n_in)
    w = F.primitive_nth_root(max(n, 2))
    got = zk.ntt(w, x, ctx)
    assert got.shape[0] == n
    assert np.array_equal(got, C.ntt(w, x))
    gi = zk.intt(w, x, ctx)
    assert np.array_equal(gi, C.ntt(w, x, inverse=True))       # n_in == 1: identity (ntt.rs:55-57)


def test_ntt_empty_input_is_an_error(ctx):
    with pytest.raises(zk.ZkbError) as e:                       # ntt.rs:11 index panic
        zk.ntt(F.primitive_nth_root(2), np.empty((0, 2), dtype=np.uint64), ctx)
    assert e.value.code == -3


def test_ntt_device_resident_and_batched(ctx):
    n, b = 1 << 13, 5
    w = F.primitive_nth_root(n)
    cols = np.stack([C.synth(100 + i, n) for i in range(b)])
    want = np.stack([C.ntt(w, cols[i]) for i in range(b)])
    got = zk.ntt_batch(w, cuda(cols.reshape(-1, 2)).reshape(b, n, 2), ctx=ctx)
    ctx.sync()
    assert np.array_equal(host(got).reshape(b, n, 2), want)
    back = zk.ntt_batch(w, got, inverse=True, ctx=ctx)
    ctx.sync()
    assert np.array_equal(host(back).reshape(b, n, 2), cols)
    assert np.array_equal(zk.ntt_batch(w, cols, ctx=ctx), want)     # host buffers, ragged tail
    rag = cols[:, :5000, :].copy()
    want_r = np.stack([C.ntt(w, rag[i]) for i in range(b)])
    assert np.array_equal(zk.ntt_batch(w, rag, ctx=ctx), want_r)


@pytest.mark.parametrize("log_n", [24, 25, 26])
def test_ntt_full_size_roundtrip_and_spot(ctx, log_n):
    """BASELINE configs[1] top size (2^24) and north_star's upper end on ONE GPU (2^25, 2^26: four register-radix passes):
    oracle on the whole vector (C oracle, seconds)."""
    n = 1 << log_n
    x = C.synth(0x5EED0005 if log_n == 26 else 0x5EED0002, n)
    w = F.primitive_nth_root(n)
    d = cuda(x)
    y = zk.ntt(w, d, ctx)
    back = zk.intt(w, y, ctx)
    ctx.sync()
    assert np.array_equal(host(back), x)
    if log_n == 26:      # the oracle's committed checksums + spot values of this transform (75 s of CPU per 2^26 oracle transform otherwise)
        from golden_ntt import check_against_golden_ntt
        check_against_golden_ntt(host(y), log_n)
    else:
        assert np.array_equal(host(y), C.ntt(w, x))


# ---------------------------------------------------------------- LDE / polynomials --------
def test_scale_kat_and_random(ctx):
    k = KATS["scale"]
    assert zk.scale(k["coeffs"], k["factor"], ctx) == k["out"]
    xs = rvals(5000)
    f = rvals(1)[0]
    assert zk.scale(xs, f, ctx) == N.scale(xs, f)


def test_device_products_at_edge_values(ctx):
    # the device Montgomery product (fe128.cuh) on operands that stress its carry / borrow paths: scale([0, a], b) = [0, a * b]
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 32, (1 << 32) - 1, 1 << 64, (1 << 96) - 1, 1 << 96, (1 << 127) - 1, 1 << 127,
            0xCB800000 << 96, (0xCB800000 << 96) - 1, 0xFFFFFFFF00000000FFFFFFFF00000000, 0x00000000FFFFFFFF00000000FFFFFFFF % P]
    for b in edge[1:]:
        xs = [0] + edge + rvals(13)
        assert zk.scale(xs, b, ctx) == [x * pow(b, i, P) % P for i, x in enumerate(xs)]
    for a in edge:
        for b in edge[1:]:
            assert zk.scale([0, a], b, ctx) == [0, a * b % P]


@pytest.mark.parametrize("log_n,n_coeffs", [(1, 1), (2, 1), (3, 2), (6, 16), (6, 13), (10, 256), (12, 1000), (12, 4096), (13, 2048),
                                            (14, 4096), (16, 1 << 14), (17, 40000), (20, 1 << 18), (21, 1 << 19), (22, 1 << 20), (22, 3)])
def test_coset_lde_vs_oracle(ctx, log_n, n_coeffs):
    n = 1 << log_n
    x = C.synth(0x5EED0003, n_coeffs)
    w = F.primitive_nth_root(n)
    want = C.coset_lde(w, n, F.GENERATOR, x)
    got = zk.fast_coset_evaluate(w, n, F.GENERATOR, x, ctx)
    assert np.array_equal(got, want)
    for g in VEC["lde"]:
        if g["log_n"] == log_n and g["n_coeffs"] == n_coeffs:
            assert sha(got) == g["sha256_out"]
            assert zk.MerkleRoot.commit(got, ctx).hex() == g["merkle_root"]


def test_coset_lde_like_reference_property_test(ctx):
    # ntt_arithmetics.rs:470-489: fast_coset_evaluate == evaluation on offset * w^k
    n = 64
    w = F.primitive_nth_root(n)
    for _ in range(5):
        a = rvals(rnd.randrange(1, 31))
        ev = zk.fast_coset_evaluate(w, n, F.GENERATOR, a, ctx)
        assert ev == N.fast_coset_evaluate(w, n, F.GENERATOR, a)
        for k in (0, 1, 17, 63):
            x = F.GENERATOR * F.fpow(w, k) % P
            assert ev[k] == sum(c * F.fpow(x, i) for i, c in enumerate(a)) % P


def test_coset_lde_edges(ctx):
    w = F.primitive_nth_root(8)
    assert zk.fast_coset_evaluate(w, 8, F.GENERATOR, [], ctx) == [0] * 8          # zero polynomial
    with pytest.raises(zk.ZkbError) as e:                                          # ntt_arithmetics.rs:168 underflow
        zk.fast_coset_evaluate(w, 8, F.GENERATOR, list(range(9)), ctx)
    assert e.value.code == -5
    assert zk.fast_coset_evaluate(1, 1, F.GENERATOR, [5], ctx) == [5]


def test_coset_lde_batch_device(ctx):
    n, nc, b = 1 << 14, 1 << 12, 6
    w = F.primitive_nth_root(n)
    cols = np.stack([C.synth(200 + i, nc) for i in range(b)])
    want = np.stack([C.coset_lde(w, n, F.GENERATOR, cols[i]) for i in range(b)])
    got = zk.coset_lde_batch(w, n, F.GENERATOR, cuda(cols.reshape(-1, 2)).reshape(b, nc, 2), ctx)
    ctx.sync()
    assert np.array_equal(host(got).reshape(b, n, 2), want)


def test_fast_multiply_and_divide_vs_schoolbook(ctx):
    # ntt_arithmetics.rs:356-380 and :491-517
    n = 64
    w = F.primitive_nth_root(n)
    for _ in range(10):
        a, b = rvals(rnd.randrange(1, 31)), rvals(rnd.randrange(1, 31))
        school = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                school[i + j] = (school[i + j] + x * y) % P
        assert zk.fast_multiply(w, n, a, b, ctx) == school == N.fast_multiply(w, n, a, b)
        assert zk.fast_coset_divide(w, n, F.GENERATOR, school, b, ctx) == a
    # trailing zero coefficients / order shrink / zero operands
    assert zk.fast_multiply(w, n, [1, 2, 0, 0], [3, 0], ctx) == N.fast_multiply(w, n, [1, 2, 0, 0], [3, 0])
    assert zk.fast_multiply(w, n, [0, 0], [3, 1], ctx) == []
    assert zk.fast_multiply(w, n, [7], [9], ctx) == [63]
    big_w = F.primitive_nth_root(1 << 14)
    a, b = rvals(5000), rvals(3000)
    assert zk.fast_multiply(big_w, 1 << 14, a, b, ctx) == N.fast_multiply(big_w, 1 << 14, a, b)
    with pytest.raises(zk.ZkbError) as e:
        zk.fast_multiply(w, 32, [1], [1], ctx)                   # root does not have that order
    assert e.value.code == -6
    with pytest.raises(zk.ZkbError) as e:
        zk.fast_coset_divide(w, n, F.GENERATOR, [1, 2], [0], ctx)
    assert e.value.code == -7
    with pytest.raises(zk.ZkbError) as e:
        zk.fast_coset_divide(w, n, F.GENERATOR, [1, 2], [1, 2, 3], ctx)
    assert e.value.code == -11
    assert zk.fast_coset_divide(w, n, F.GENERATOR, [0], [1, 2, 3], ctx) == []


# ---------------------------------------------------------------- Merkle -------------------
def test_merkle_reference_kats(ctx):
    for k in KATS["merkle"]["commit"]:
        assert zk.MerkleRoot.commit(k["leafs"], ctx).hex() == k["root"]
    o = KATS["merkle"]["open"]
    path = zk.MerkleRoot.open(o["index"], o["leafs"], ctx)
    assert [h.hex() for h in path] == o["path"]
    root = bytes.fromhex(KATS["merkle"]["commit"][-1]["root"])
    assert zk.MerkleRoot.verify(root, 1, path, 456)
    assert not zk.MerkleRoot.verify(root, 1, path, 5462)
    assert not zk.MerkleRoot.verify(root, 0, path, 456)


@pytest.mark.parametrize("log_n", list(range(0, 17)) + [20, 21])
def test_merkle_commit_vs_oracle(ctx, log_n):
    n = 1 << log_n
    vals = C.synth(0x5EED0004, n)
    want = C.merkle(vals)
    assert zk.MerkleRoot.commit(vals, ctx) == want
    assert zk.MerkleRoot.commit(cuda(vals), ctx) == want
    for g in VEC["merkle"]:
        if g["log_n"] == log_n:
            assert want.hex() == g["root"]


def test_merkle_leaf_encoding_edges(ctx):
    """decimal-ASCII leaf preimages of every length 1..39 (field_element.rs:46-50)"""
    vals = [0, 1, 9] + [10 ** k for k in range(1, 39)] + [10 ** k - 1 for k in range(1, 39)] + [P - 1, P - 2, 1 << 64, (1 << 64) - 1,
            (1 << 96) - 1, 1 << 96, 1 << 127, 5462, 11]
    vals = (vals + [3] * 128)[:128]
    assert zk.MerkleRoot.commit(vals, ctx) == M.commit(vals)
    big = (vals * 16)[:2048]                    # through the 1024-leaf tile kernel as well
    assert zk.MerkleRoot.commit(big, ctx) == C.merkle(C.to_arr(big))


def test_merkle_not_power_of_two(ctx):
    for n in (0, 3, 6, 1000):
        with pytest.raises(zk.ZkbError) as e:                    # merkle_root.rs:9
            zk.MerkleRoot.commit(C.synth(1, n) if n else np.empty((0, 2), dtype=np.uint64), ctx)
        assert e.value.code == -4


@pytest.mark.parametrize("log_n", [1, 2, 5, 6, 10, 11, 12, 16, 20])
def test_merkle_open_vs_oracle(ctx, log_n):
    n = 1 << log_n
    vals = C.synth(0x5EED0005, n)
    t = fastfri.Tree(vals)
    tree = zk.MerkleTree(cuda(vals), ctx)
    assert tree.root() == t.root
    idx = list(range(n)) if n <= 64 else [0, 1, 31, 32, 33, n // 2 - 1, n // 2, n - 2, n - 1] + [rnd.randrange(n) for _ in range(40)]
    paths = tree.open_many(idx)
    vl = C.from_arr(vals[idx])
    for i, p, v in zip(idx, paths, vl):
        assert p == t.open(i)
        assert zk.MerkleRoot.verify(t.root, i, p, v)
        assert M.verify(t.root, i, p, v)
    with pytest.raises(zk.ZkbError) as e:
        tree.open(n)
    assert e.value.code == -8
    tree.close()


# ---------------------------------------------------------------- FRI ----------------------
@pytest.mark.parametrize("log_n", [1, 2, 5, 10, 11, 12, 15, 20])
def test_fri_fold_vs_oracle(ctx, log_n):
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    cw = C.synth(0x5EED0006, n)
    alpha = rvals(1)[0]
    fri = zk.FRI(F.GENERATOR, w, n, 4, 1, ctx)
    got = fri.fold(cw, alpha)
    assert np.array_equal(got, C.fri_fold(cw, alpha, F.GENERATOR, w))
    if n <= 1 << 10:
        assert C.from_arr(got) == OFRI.fold(C.from_arr(cw), alpha, F.GENERATOR, w)      # literal fri.rs:150-159
    off2 = F.mul(F.GENERATOR, F.GENERATOR)
    assert np.array_equal(fri.fold(cw, 0, off2, w), C.fri_fold(cw, 0, off2, w))


def _codeword(log_n, seed=0x5EED0003):
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    return n, w, C.coset_lde(w, n, F.GENERATOR, C.synth(seed, n // 4))


def test_fri_prove_like_reference_test(ctx):
    # src/fri.rs:451-531: degree 63, N = 256, ef 4, 17 colinearity tests
    degree, ef, ncc = 63, 4, 17
    n = (degree + 1) * ef
    w = F.primitive_nth_root(n)
    poly = list(range(degree + 1))
    codeword = zk.ntt(w, poly + [0] * (n - len(poly)), ctx)
    assert codeword == N.ntt(w, poly + [0] * (n - len(poly)))
    fri = zk.FRI(F.GENERATOR, w, n, ef, ncc, ctx)
    ofri = OFRI(F.GENERATOR, w, n, ef, ncc)
    ps = zk.IndependentProofStream()
    top = fri.prove(codeword, ps)
    ops = PS.IndependentProofStream()
    assert top == ofri.prove(codeword, ops)
    assert ps.digest() == ops.digest()                      # final proof-stream bytes, bit-exact
    points = []
    assert ofri.verify(PS.IndependentProofStream(PS.parse(ps.digest())), points) is None
    for x, y in points:
        assert sum(c * F.fpow(w, x * i) for i, c in enumerate(poly)) % P == y
    # disturbed codeword must be rejected by the (restated) reference verifier (fri.rs:514-528)
    bad = [0] * (degree // 3) + codeword[degree // 3:]
    ps = zk.IndependentProofStream()
    fri.prove(bad, ps)
    assert ofri.verify(PS.IndependentProofStream(PS.parse(ps.digest())), []) is not None


@pytest.mark.parametrize("case", VEC["fri"], ids=lambda c: "2^%d%s" % (c["log_n"], "-sig" if c["document"] else ""))
def test_fri_prove_vs_oracle_and_golden(ctx, case):
    log_n, ncc, doc = case["log_n"], case["ncc"], case["document"]
    n, w, cw = _codeword(log_n, case["seed"])
    fri = zk.FRI(F.GENERATOR, w, n, 4, ncc, ctx)
    ofri = OFRI(F.GENERATOR, w, n, 4, ncc)
    # (1) one-call C path (zkb_fri_prove) with the library's own proof stream
    ps = zk.SignatureProofStream(doc.encode()) if doc else zk.IndependentProofStream()
    top = fri.prove(cuda(cw), ps)
    d = ps.digest()
    assert top == case["top_indices"]
    assert len(d) == case["proof_bytes"]
    assert hashlib.sha256(d).hexdigest() == case["sha256_proof"]
    # (2) commit/query through the Fiat-Shamir callback with a foreign (oracle) proof stream
    ops = PS.SignatureProofStream(doc.encode()) if doc else PS.IndependentProofStream()
    layers = fri.commit(cw, ops)
    assert [layers.root(r).hex() for r in range(len(layers))] == case["roots"]
    layers.close()
    if log_n <= 16:
        # (3) against the oracle run live, layer by layer
        ops2 = PS.SignatureProofStream(doc.encode()) if doc else PS.IndependentProofStream()
        top2, codewords, trees = fastfri.prove(ofri, cw, ops2)
        assert top2 == top and ops2.digest() == d
        ops3 = PS.SignatureProofStream(doc.encode()) if doc else PS.IndependentProofStream()
        layers = fri.commit(cuda(cw), ops3)
        for r in range(len(layers)):
            assert np.array_equal(layers.codeword(r), codewords[r])
        layers.close()
        vps = PS.SignatureProofStream(doc.encode(), PS.parse(d)) if doc else PS.IndependentProofStream(PS.parse(d))
        assert ofri.verify(vps, []) is None


def test_fri_python_stream_path_equals_c_path(ctx):
    n, w, cw = _codeword(13)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    a = zk.IndependentProofStream()
    b = zk.IndependentProofStream()
    b.force_python = True
    assert fri.prove(cw, a) == fri.prove(cw, b)
    assert a.digest() == b.digest()


def test_fri_errors(ctx):
    n, w, cw = _codeword(10)
    with pytest.raises(AssertionError):                          # fri.rs:215-219
        zk.FRI(F.GENERATOR, w, n * 2, 4, 64, ctx).prove(cw, zk.IndependentProofStream())
    with pytest.raises(zk.ZkbError) as e:                        # fri.rs:133
        zk.FRI(F.GENERATOR, F.mul(w, 3), n, 4, 64, ctx).prove(cw, zk.IndependentProofStream())
    assert e.value.code == -6
    with pytest.raises(zk.ZkbError) as e:                        # fri.rs:225: < 2 rounds
        zk.FRI(F.GENERATOR, F.primitive_nth_root(256), 256, 4, 64, ctx).prove(cw[:256], zk.IndependentProofStream())
    assert e.value.code == -10


def test_lde_fri_commit_full_size_2_24(ctx):
    """BASELINE configs[2]: coset LDE + Merkle + full FRI commit on one 2^24 codeword.
    Checked (a) against the C oracle for the codeword, every root and every layer, and
    (b) by the restated reference verifier accepting the proof bytes."""
    log_n = 24
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    coeffs = C.synth(0x5EED0003, n // 4)
    cw = zk.fast_coset_evaluate(w, n, F.GENERATOR, cuda(coeffs), ctx)
    ctx.sync()
    want_cw = C.coset_lde(w, n, F.GENERATOR, coeffs)
    assert np.array_equal(host(cw), want_cw)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    ofri = OFRI(F.GENERATOR, w, n, 4, 64)
    assert fri.num_rounds() == 16
    ps = zk.IndependentProofStream()
    top = fri.prove(cw, ps)
    d = ps.digest()
    ops = PS.IndependentProofStream()
    top2, codewords, trees = fastfri.prove(ofri, want_cw, ops)
    assert top == top2
    assert d == ops.digest()
    assert ofri.verify(PS.IndependentProofStream(PS.parse(d)), []) is None


@pytest.mark.parametrize("log_n,n_coeffs", [(10, 256), (12, 1000), (16, 1 << 14), (16, 0)])
def test_lde_fri_commit_pipeline(ctx, log_n, n_coeffs):
    """zkb_lde_fri_commit: LDE -> FRI commit without leaving HBM == the two calls separately."""
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    coeffs = C.synth(0x5EED0003, n_coeffs) if n_coeffs else np.empty((0, 2), dtype=np.uint64)
    cw = C.coset_lde(w, n, F.GENERATOR, coeffs) if n_coeffs else np.zeros((n, 2), dtype=np.uint64)
    ofri = OFRI(F.GENERATOR, w, n, 4, 64)
    ops = PS.IndependentProofStream()
    codewords, trees, _ = fastfri.commit(ofri, cw, ops)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    for stream in (zk.IndependentProofStream(), PS.IndependentProofStream()):      # C callback / Python callback
        layers = fri.lde_commit(cuda(coeffs) if n_coeffs else coeffs, stream)
        assert stream.digest() == ops.digest()
        assert [layers.root(r) for r in range(len(layers))] == [t.root for t in trees]
        for r in range(len(layers)):
            assert np.array_equal(layers.codeword(r), codewords[r])
        layers.close()


def test_profile_counters(ctx):
    ctx.profile(True, reset=True)
    n = 1 << 20
    zk.MerkleRoot.commit(C.synth(1, n), ctx)
    prof = ctx.profile_read()
    ctx.profile(False)
    assert prof["k_leaf8<false>"][1] == 1 and prof["k_leaf8<false>"][0] > 0
    assert prof["k_tree"][1] == 1


def test_fast_multiply_trailing_zero_quirk(ctx):
    """Operands with trailing zero coefficients longer than the shrunk transform order: the
    reference's ntt() then runs at a longer length with a root of too small an order
    (ntt_arithmetics.rs:38-47) - a reference quirk that the oracle restates literally and the
    CUDA path must reproduce bit for bit."""
    n = 64
    w = F.primitive_nth_root(n)
    cases = [([1, 2, 0, 0], [3, 0]), ([5, 6, 7, 0, 0, 0, 0, 0, 0], [1, 1]), ([7, 0], [9, 0, 0]), ([1, 2, 3], [4, 5, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])]
    for _ in range(5):
        a, b = rvals(rnd.randrange(1, 8)), rvals(rnd.randrange(1, 8))
        cases.append((a + [0] * rnd.randrange(1, 40), b + [0] * rnd.randrange(0, 40)))
    for a, b in cases:
        assert zk.fast_multiply(w, n, a, b, ctx) == N.fast_multiply(w, n, a, b)
    for a, b in cases:
        prod = N.fast_multiply(w, n, [x for x in a if x], [x for x in b if x] or [1])
        bb = [x for x in b if x] or [1]
        assert zk.fast_coset_divide(w, n, F.GENERATOR, prod + [0] * 9, bb + [0] * 5, ctx) == \
            N.fast_coset_divide(w, n, F.GENERATOR, prod + [0] * 9, bb + [0] * 5)


def test_fast_multiply_quirk_beyond_one_tile_and_device_operands(ctx):
    """The same quirk with operands longer than one 4096-value tile (the literal loop then runs as global-memory stages) and
    with operands / results that live in device memory: no size or residency limit versus the reference's signatures."""
    n = 1 << 14
    w = F.primitive_nth_root(n)
    a, b = rvals(30) + [0] * 5000, rvals(9) + [0] * 6000            # degrees 29 + 8: order shrinks to 64, operands pad to 8192
    want = N.fast_multiply(w, n, a, b)
    assert zk.fast_multiply(w, n, a, b, ctx) == want
    prod = want + [0] * 4500
    assert zk.fast_coset_divide(w, n, F.GENERATOR, prod, b, ctx) == N.fast_coset_divide(w, n, F.GENERATOR, prod, b)
    # device-resident operands through the C ABI
    import ctypes
    from zk_stark_tutor_b200.context import le16, pack, unpack
    x, y = rvals(3000), rvals(2500)
    dx, dy = cuda(pack(x)), cuda(pack(y))
    import torch
    dout = torch.empty((len(x) + len(y), 2), dtype=torch.int64, device="cuda")
    n_out = ctypes.c_size_t(0)
    ctx.check(ctx.lib.zkb_poly_mul(ctx.h, le16(w), n, dx.data_ptr(), len(x), dy.data_ptr(), len(y), dout.data_ptr(), ctypes.byref(n_out)))
    assert unpack(host(dout)[:n_out.value]) == zk.fast_multiply(w, n, x, y, ctx)


def test_ntt_with_non_primitive_root_matches_reference_loop(ctx):
    """ntt() does not check its root (ntt.rs:7-49); for lengths up to one tile the CUDA path
    is the same radix-2 DIT op for op, so even a wrong-order root gives the reference's result."""
    for n, order in ((8, 4), (16, 2), (64, 16), (1024, 256), (4096, 64)):
        r = F.primitive_nth_root(order)
        xs = rvals(n)
        assert zk.ntt(r, xs, ctx) == N.ntt(r, xs)


def test_columns_lde_commit_matches_oracle(ctx):
    """BASELINE configs[3] in miniature: independent columns -> LDE + FRI commit each; roots
    gathered (world = 1 here; the 2/3-rank gather is covered on CPU by test_columns_gloo.py)."""
    from zk_stark_tutor_b200 import columns
    n, n_cols = 1 << 13, 5
    w = F.primitive_nth_root(n)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    ofri = OFRI(F.GENERATOR, w, n, 4, 64)
    mine = columns.partition(n_cols, 1, 0)
    coeffs = [C.synth(0x5EED0004 + c, n // 4) for c in mine]
    got = columns.lde_commit_columns(fri, [cuda(x) for x in coeffs], zk.IndependentProofStream, ctx)
    allr = columns.gather_roots([g[0] for g in got], n_cols, 1, 0, fri.num_rounds())
    for c, x in zip(mine, coeffs):
        ops = PS.IndependentProofStream()
        _, trees, _ = fastfri.commit(ofri, C.coset_lde(w, n, F.GENERATOR, x), ops)
        assert got[c][1] == ops.digest()
        assert [bytes(allr[c, r]) for r in range(len(trees))] == [t.root for t in trees]
    # the same columns with several in flight on one GPU (3 streams, 3 host threads)
    import torch
    torch.cuda.synchronize()
    pipe = columns.ColumnPipeline(0, (F.GENERATOR, w, n, 4, 64), lanes=3)
    got2 = pipe.run([cuda(x) for x in coeffs], zk.IndependentProofStream)
    pipe.close()
    assert [g[1] for g in got2] == [g[1] for g in got] and [g[0] for g in got2] == [g[0] for g in got]


# ---------------------------------------------------------------- a10: domain algorithms ----
def test_domain_algorithms_vs_oracle(ctx):
    """fast_zerofier / fast_evaluate_domain / fast_interpolate_domain (ntt_arithmetics.rs:66-237;
    reference property tests :382-468) - NTT products on the GPU, scalar glue on the host."""
    from oracle import poly as PL
    n = 64
    w = F.primitive_nth_root(n)
    for size in (0, 1, 2, 5, 16, 23):
        domain = rvals(size)
        z = zk.fast_zerofier(w, n, domain, ctx)
        assert z == PL.fast_zerofier(w, n, domain) == (PL.zerofier_domain(domain) if size else [])
        poly = rvals(rnd.randrange(1, 31))
        ev = zk.fast_evaluate_domain(w, n, poly, domain, ctx)
        assert ev == PL.fast_evaluate_domain(w, n, poly, domain) == [PL.evaluate(poly, x) for x in domain]
        values = rvals(size)
        interp = zk.fast_interpolate_domain(w, n, domain, values, ctx)
        assert interp == PL.fast_interpolate_domain(w, n, domain, values)
        assert [PL.evaluate(interp, x) for x in domain] == values
    # the Stark prover's use (stark.rs:305-326): interpolate a 284-row trace column on omicron^i
    n = 1024
    w = F.primitive_nth_root(n)
    domain = [F.fpow(w, i) for i in range(284)]
    values = rvals(284)
    interp = zk.fast_interpolate_domain(w, n, domain, values, ctx)
    assert interp == PL.fast_interpolate_domain(w, n, domain, values)
    with pytest.raises(zk.ZkbError) as e:
        zk.fast_zerofier(w, 512, domain, ctx)
    assert e.value.code == -6


# ---------------------------------------------------------------- the caller's sequence -----
def _stark_like_prove(B, seed, n_regs=2, omicron_len=1024, ef=4, ncc=64, trace_len=284, tc_degree=3):
    """The hot-path call sequence of Stark::prove (stark.rs:363-562) with synthetic quotient
    polynomials in place of the Rescue-Prime AIR (which is out of scope): per register LDE +
    Merkle commit, randomizer LDE + commit, Fiat-Shamir weights (incl. the reference's
    all-weights-equal quirk, stark.rs:262-274), x^shift products via fast_multiply, weighted
    combination, LDE, FRI::prove, quadrupled sorted indices, Value + Path openings.
    `B` bundles the backend functions; the same function runs on the oracle and on the GPU."""
    r = random.Random(seed)
    fri_len = omicron_len * ef
    omega, omicron = F.primitive_nth_root(fri_len), F.primitive_nth_root(omicron_len)
    g = F.GENERATOR
    ps = B["stream"]()
    fri = B["fri"](g, omega, fri_len, ef, ncc)
    bq = [[r.randrange(P) for _ in range(trace_len - 2)] for _ in range(n_regs)]          # boundary quotients
    max_deg = (1 << (tc_degree * (trace_len - 1)).bit_length()) - 1                         # stark.rs:186-200
    tq_bound = tc_degree * (trace_len - 1) - (trace_len - 256 - 1)
    tq = [[r.randrange(P) for _ in range(tq_bound + 1)] for _ in range(n_regs)]             # transition quotients
    bq_cw = []
    for s in range(n_regs):
        cw = B["lde"](omega, fri_len, g, bq[s])
        ps.push((PS.ROOT, B["commit"](cw)))
        bq_cw.append(cw)
    randomizer = [F.sample(bytes(r.randrange(256) for _ in range(17))) for _ in range(max_deg + 1)]
    r_cw = B["lde"](omega, fri_len, g, randomizer)
    ps.push((PS.ROOT, B["commit"](r_cw)))
    randomness = ps.fiat_shamir_prover(32)
    weights = [F.sample(bytes(i) + randomness) for i in range(1 + 2 * len(tq) + 2 * len(bq))]
    assert len(set(weights)) == 1
    x_pow = lambda k: [0] * k + [1]
    terms = [randomizer]
    for t in tq:
        terms += [t, B["mul"](omicron, omicron_len, x_pow(max_deg - tq_bound), t)]
    for b in bq:
        terms += [b, B["mul"](omicron, omicron_len, x_pow(max_deg - (len(b) - 1)), b)]
    from oracle import poly as PL
    comb = []
    for wgt, term in zip(weights, terms):
        comb = PL.add(comb, PL.mul([wgt], term)) if comb else PL.mul([wgt], term)
    comb_cw = B["lde"](omega, fri_len, g, comb)
    idx = B["prove"](fri, comb_cw, ps)
    dup = idx + [(i + ef) % fri_len for i in idx]
    quad = sorted(dup + [(i + fri_len // 2) % fri_len for i in dup])
    for cw in bq_cw + [r_cw]:
        if "open_into" in B:                          # the same loop as one C call (zkb_merkle_open_ps)
            B["open_into"](cw, quad, ps)
            continue
        paths = B["open_many"](cw, quad)
        for i, path in zip(quad, paths):
            ps.push((PS.VALUE, cw[i]))
            ps.push((PS.PATH, path))
    return ps.digest()


def test_stark_call_sequence_proof_bytes(ctx):
    """GPU-backed and oracle-backed runs of the Stark prover's hot-path sequence produce the same
    proof bytes, of exactly the size the reference quotes for these shapes (src/rpsss.rs:89)."""
    def oracle_open_many(cw, idxs):
        t = fastfri.Tree(C.to_arr(cw))
        return [t.open(i) for i in idxs]

    def oracle_prove(fri, cw, ps):
        top, _, _ = fastfri.prove(fri, C.to_arr(cw), ps)
        return top

    oracle = {"stream": PS.IndependentProofStream, "fri": OFRI,
              "lde": N.fast_coset_evaluate, "commit": M.commit, "mul": N.fast_multiply,
              "prove": oracle_prove, "open_many": oracle_open_many}

    def gpu_open_many(cw, idxs):
        t = zk.MerkleTree(cw, ctx)
        try:
            return t.open_many(idxs)
        finally:
            t.close()

    gpu = {"stream": zk.IndependentProofStream, "fri": lambda *a: zk.FRI(*a, ctx=ctx),
           "lde": lambda w, n, off, p: zk.fast_coset_evaluate(w, n, off, p, ctx),
           "commit": lambda cw: zk.MerkleRoot.commit(cw, ctx),
           "mul": lambda w, n, a, b: zk.fast_multiply(w, n, a, b, ctx),
           "prove": lambda fri, cw, ps: fri.prove(cw, ps), "open_many": gpu_open_many}
    want = _stark_like_prove(oracle, 4242)
    got = _stark_like_prove(gpu, 4242)
    assert len(want) == 1156888                                   # src/rpsss.rs:89
    assert got == want
    # the opening loop as one batched C call appending Value/Path objects itself
    def gpu_open_into(cw, idxs, ps):
        t = zk.MerkleTree(cw, ctx)
        try:
            t.open_into(idxs, ps)
        finally:
            t.close()
    assert _stark_like_prove(dict(gpu, open_into=gpu_open_into), 4242) == want
    # the signature flavour (document-prefixed Fiat-Shamir, rescue_prime/proof_stream.rs)
    oracle["stream"] = lambda: PS.SignatureProofStream(b"a document")
    gpu["stream"] = lambda: zk.SignatureProofStream(b"a document")
    assert _stark_like_prove(gpu, 7) == _stark_like_prove(oracle, 7)


# ---------------------------------------------------------------- RPSSS-shaped proof batch ---
def _oracle_hot_path(shape, columns, combination, ps):
    """proofs.prove_hot_path restated on the oracle (same inputs, same transcript)."""
    n = shape.fri_len
    w = F.primitive_nth_root(n)
    cws = []
    for col in columns:
        cw = C.coset_lde(w, n, F.GENERATOR, col)
        cws.append(cw)
        ps.push((PS.ROOT, fastfri.Tree(cw).root))
    fri = OFRI(F.GENERATOR, w, n, shape.ef, shape.ncc)
    top, _, _ = fastfri.prove(fri, C.coset_lde(w, n, F.GENERATOR, combination), ps)
    from zk_stark_tutor_b200.proofs import quadrupled_indices
    quad = quadrupled_indices(top, n, shape.ef)
    for cw in cws:
        t = fastfri.Tree(cw)
        vals = C.from_arr(cw)
        for i in quad:
            ps.push((PS.VALUE, vals[i]))
            ps.push((PS.PATH, t.open(i)))
    return top


def test_rpsss_shaped_proof_batch(ctx):
    """BASELINE configs[4], second half: a batch of signature-shaped proofs through the native call
    sequence, several in flight per GPU; every proof's bytes == the oracle's, 1,156,888 bytes each."""
    from zk_stark_tutor_b200 import proofs
    shape = proofs.ProofShape()
    n_proofs = 5
    inputs = []
    for p in range(n_proofs):
        cols = [C.synth(0x5EED0005 + 16 * p + k, ln) for k, ln in enumerate(shape.column_lengths())]
        inputs.append((cols, C.synth(0x5EED0005 + 16 * p + 15, shape.comb_len)))
    want = []
    for cols, comb in inputs[:2]:
        ops = PS.SignatureProofStream(b"a document")
        top = _oracle_hot_path(shape, cols, comb, ops)
        want.append((top, ops.digest()))
    assert len(want[0][1]) == 1156888                                   # src/rpsss.rs:89
    w = F.primitive_nth_root(shape.fri_len)
    fri = zk.FRI(F.GENERATOR, w, shape.fri_len, shape.ef, shape.ncc, ctx)
    for (cols, comb), (top, digest) in zip(inputs[:2], want):
        ps = zk.SignatureProofStream(b"a document")
        assert proofs.prove_hot_path(ctx, fri, shape, cols, comb, ps) == top
        assert ps.digest() == digest
    # the whole batch, 3 lanes, device-resident inputs; lanes must not disturb each other
    import torch
    torch.cuda.synchronize()
    pipe = proofs.ProofPipeline(0, shape, F.GENERATOR, w, lanes=3)
    dev_inputs = [([cuda(c) for c in cols], cuda(comb)) for cols, comb in inputs]
    got = pipe.run(dev_inputs, lambda: zk.SignatureProofStream(b"a document"), keep_digest=True)
    again = pipe.run(dev_inputs, lambda: zk.SignatureProofStream(b"a document"), keep_digest=True)
    pipe.close()
    assert [g[0] for g in got] == [1156888] * n_proofs
    assert got[0][1] == want[0][1] and got[1][1] == want[1][1]
    assert [g[1] for g in got] == [g[1] for g in again]
    assert len({g[1] for g in got}) == n_proofs


# ---------------------------------------------------------------- the real callers ----------
def _gpu_backend(ctx):
    """The hot-path functions Stark / RPSSS call (oracle.stark.Backend), served by the CUDA path
    through the mirror of the reference's API."""
    class GpuBackend:
        name = "gpu"
        fast_zerofier = staticmethod(lambda root, order, domain: zk.fast_zerofier(root, order, domain, ctx))
        fast_interpolate_domain = staticmethod(lambda root, order, dom, vals: zk.fast_interpolate_domain(root, order, dom, vals, ctx))
        fast_coset_divide = staticmethod(lambda root, order, off, a, b: zk.fast_coset_divide(root, order, off, a, b, ctx))
        fast_coset_evaluate = staticmethod(lambda w, n, off, p: zk.fast_coset_evaluate(w, n, off, p, ctx))
        fast_multiply = staticmethod(lambda root, order, a, b: zk.fast_multiply(root, order, a, b, ctx))
        commit = staticmethod(lambda cw: zk.MerkleRoot.commit(cw, ctx))

        @staticmethod
        def open_many(cw, idxs):
            t = zk.MerkleTree(cw, ctx)
            try:
                return t.open_many(idxs)
            finally:
                t.close()

        @staticmethod
        def fri(offset, omega, n, ef, ncc):
            return zk.FRI(offset, omega, n, ef, ncc, ctx=ctx)
    return GpuBackend


def test_rpsss_signature_gpu_backend(ctx):
    """BASELINE configs[0] / configs[4]: the Rescue-Prime STARK signature (src/rpsss.rs:70-87 ->
    Stark::prove, stark.rs:276-563) with every hot-path call - interpolation, zerofiers, coset
    division, the products inside the AIR evaluation, LDEs, Merkle commits and openings, FRI::prove -
    served by the CUDA path: the 1,156,888-byte signature is byte-identical to the oracle-backed one
    and the restated Stark::verify (stark.rs:565-770) accepts it for the right document only."""
    from oracle.stark import RPSSS, deterministic_rng
    cpu = RPSSS(4, 64, 128, 3)
    gpu = RPSSS(4, 64, 128, 3, backend=_gpu_backend(ctx))
    assert [t.dictionary for t in gpu.transition_constraints()] == [t.dictionary for t in cpu.transition_constraints()]
    sk, pk = cpu.keygen(deterministic_rng(b"k"))
    doc = b"Hello, World!"
    want = cpu.sign(sk, doc, deterministic_rng(b"r"))
    got = gpu.sign(sk, doc, deterministic_rng(b"r"), make_stream=zk.SignatureProofStream)
    assert len(got) == 1156888
    assert got == want
    assert cpu.verify(pk, doc, got) is None
    assert cpu.verify(pk, b"Malicious document", got) is not None


def test_stark_prove_gpu_backend_independent_stream(ctx):
    """Stark::prove on a Rescue-Prime hash trace over an IndependentProofStream (stark.rs:810-880), GPU-backed."""
    from oracle.rescue_prime import RescuePrime
    from oracle.stark import Stark, deterministic_rng
    B = _gpu_backend(ctx)
    rp = RescuePrime(2, 1, 128, 27, interpolate=B.fast_interpolate_domain)
    x = 0xFEEDFACE12345
    out = rp.hash(x)
    got_stark = Stark(4, 64, 128, rp.m, rp.N + 1, 3, backend=B)
    ref_stark = Stark(4, 64, 128, rp.m, rp.N + 1, 3)
    tcs = rp.transition_constraints(ref_stark.omicron, ref_stark.omicron_domain_length)
    got = got_stark.prove(rp.trace(x), tcs, rp.boundary_constraints(out), zk.IndependentProofStream(), deterministic_rng(b"z"))
    want = ref_stark.prove(rp.trace(x), tcs, rp.boundary_constraints(out), PS.IndependentProofStream(), deterministic_rng(b"z"))
    assert got == want
    assert ref_stark.verify(tcs, rp.boundary_constraints(out), PS.IndependentProofStream(PS.parse(got))) is None


def test_rpsss_shaped_proofs_in_lockstep(ctx):
    """csrc/batch.cu: B proofs advance together (every launch carries all instances); each proof's bytes are
    identical to the one-at-a-time sequence and to the oracle."""
    from zk_stark_tutor_b200 import proofs
    shape = proofs.ProofShape()
    B = 7
    inputs = []
    for p in range(B):
        cols = [C.synth(0x5EED0005 + 16 * p + k, ln) for k, ln in enumerate(shape.column_lengths())]
        inputs.append((cols, C.synth(0x5EED0005 + 16 * p + 15, shape.comb_len)))
    w = F.primitive_nth_root(shape.fri_len)
    fri = zk.FRI(F.GENERATOR, w, shape.fri_len, shape.ef, shape.ncc, ctx)
    one_by_one = []
    for cols, comb in inputs:
        ps = zk.SignatureProofStream(b"a document")
        top = proofs.prove_hot_path(ctx, fri, shape, cols, comb, ps)
        one_by_one.append((top, ps.digest()))
    ops = PS.SignatureProofStream(b"a document")
    _oracle_hot_path(shape, inputs[0][0], inputs[0][1], ops)
    assert one_by_one[0][1] == ops.digest()
    packed = proofs.pack_batch(shape, inputs)
    for src in (packed, cuda(packed.reshape(-1, 2)).reshape(packed.shape)):
        streams = [zk.SignatureProofStream(b"a document") for _ in range(B)]
        tops = proofs.prove_hot_path_batch(ctx, fri, shape, src, streams)
        assert tops == [t for t, _ in one_by_one]
        assert [s.digest() for s in streams] == [d for _, d in one_by_one]
    # a batch of one, and batched trees / openings against the single-tree calls on ragged sizes
    streams = [zk.SignatureProofStream(b"a document")]
    assert proofs.prove_hot_path_batch(ctx, fri, shape, proofs.pack_batch(shape, inputs[:1]), streams) == [one_by_one[0][0]]
    assert streams[0].digest() == one_by_one[0][1]
