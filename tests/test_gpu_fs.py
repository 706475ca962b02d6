"""GPU parity of the device Fiat-Shamir path: SHAKE256 by the warp sponge (csrc/keccak.cuh), the tree kernels that
draw the challenge on the device, and the persistent tail kernel (csrc/fri_tail.cu) - against the oracle's literal
restatement of FRI::commit (src/fri.rs:115-172) over proof_stream.rs:36-40 / proof_stream_enum.rs:67-190, bit-exact."""
import ctypes
import hashlib
import os
import random

import numpy as np
import pytest

import zk_stark_tutor_b200 as zk
from zk_stark_tutor_b200 import proof_stream as ZPS
from oracle import cbind as C, field as F, proof_stream as PS, fastfri
from oracle.fri import FRI as OFRI

pytestmark = pytest.mark.gpu
rnd = random.Random(2024)


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def cuda(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def _buf(b):
    return (ctypes.c_uint8 * max(len(b), 1)).from_buffer_copy(bytes(b) or b"\0")


def test_shake256_device_kats(ctx):
    # src/proof_stream.rs:129-145 (the stream there digests as its Debug string), on the device sponge
    cases = [(b"[]", "ec784925b52067bce01fd820f554a34a3f8522b337f82e00ea03d3fa2b207ef9c2c1b9ed900cf2bbfcd19a232a94c6121e041615305c4155d46d52f58a8cff1c"),
             (b'[Str("Hello, World!"), Vec([0, 1, 5, 234]), Map({"something": 123})]',
              "78b0db5cfd13c78498fd0951a9fd609f2521fd02d850cc561eced844bb0c338588358abcc0d98d76c6779cb388514f4bc19e2c0125b143abee166cb98c38a831")]
    for msg, want in cases:
        out = (ctypes.c_uint8 * 64)()
        ctx.check(ctx.lib.zkb_shake256_device(ctx.h, _buf(msg), len(msg), out, 64))
        assert bytes(out).hex() == want
    for n in (0, 1, 7, 8, 72, 73, 135, 136, 137, 271, 272, 273, 1000, 5000):
        msg = bytes(rnd.randrange(256) for _ in range(n))
        for olen in (32, 136):
            out = (ctypes.c_uint8 * olen)()
            ctx.check(ctx.lib.zkb_shake256_device(ctx.h, _buf(msg) if n else None, n, out, olen))
            assert bytes(out) == hashlib.shake_256(msg).digest(olen), (n, olen)


def _codeword(log_n, seed=0x5EED0003):
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    return n, w, C.coset_lde(w, n, F.GENERATOR, C.synth(seed, n // 4))


def _prefill(stream, objs):
    for o in objs:
        stream.push(o)


# transcripts that exist BEFORE FRI::commit starts: what Stark::prove has pushed by then (roots only: the zero-header
# quirk of proof_stream_enum.rs:186-188; or a Value: order header), at lengths that move the 136-byte block boundary
# across every position of the 73-byte Root records
PREFILLS = [
    [],
    [(PS.ROOT, bytes(range(64)))],
    [(PS.ROOT, bytes(range(64))), (PS.ROOT, bytes(range(1, 65)))],
    [(PS.ROOT, bytes(range(64)))] * 5,
    [(PS.VALUE, 12345678901234567890123456789)],
    [(PS.ROOT, bytes(7)), (PS.VALUE, 3), (PS.ROOT, bytes(63))],
    [(PS.ROOT, bytes(k)) for k in range(1, 40)],
]


@pytest.mark.parametrize("log_n", [10, 11, 13, 16, 17, 18, 19])
@pytest.mark.parametrize("pre", range(len(PREFILLS)))
def test_device_fs_commit_vs_oracle(ctx, log_n, pre):
    """zkb_fri_commit_ps (device Fiat-Shamir, persistent tail) == the oracle's FRI::commit, for Independent and
    Signature streams with transcripts of different lengths before the commit: roots, every layer, stream bytes."""
    if log_n > 13 and pre not in (0, 3, 5):
        pytest.skip("long prefills are exercised on the small sizes")
    n, w, cw = _codeword(log_n, 0x5EED0003 + pre)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    ofri = OFRI(F.GENERATOR, w, n, 4, 64)
    for doc in (None, b"a document to sign"):
        ps = zk.SignatureProofStream(doc) if doc else zk.IndependentProofStream()
        ops = PS.SignatureProofStream(doc) if doc else PS.IndependentProofStream()
        _prefill(ps, PREFILLS[pre]); _prefill(ops, PREFILLS[pre])
        codewords, trees, _ = fastfri.commit(ofri, cw, ops)
        layers = fri.commit(cuda(cw), ps)
        assert ps.digest() == ops.digest()
        assert [layers.root(r) for r in range(len(layers))] == [t.root for t in trees]
        for r in range(len(layers)):
            assert np.array_equal(layers.codeword(r), codewords[r]), r
        # the host stream continues exactly where the device sponge stopped
        assert ps.fiat_shamir_prover(32) == ops.fiat_shamir_prover(32)
        layers.close()


@pytest.mark.parametrize("log_n", [12, 16, 18, 20])
def test_device_fs_equals_host_hop_path(ctx, log_n):
    """Same call with ZKB_HOST_FS=1 (one host hop per round, the foreign-stream path) and with ZKB_HOST_ASSEMBLY=1 (query-phase objects
    framed on the host from raw paths instead of by k_open_wire / k_leafs_wire; pruned trees above 2^17 leaves): identical proof bytes,
    also for the Value + Path openings of zkb_merkle_open_ps."""
    n, w, cw = _codeword(log_n)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    idx = (ctypes.c_uint64 * 9)(0, 1, n - 1, n // 2, n // 2 - 1, 12345 % n, 777 % n, n // 3, 5)
    d = []
    for env in ({}, {"ZKB_HOST_FS": "1"}, {"ZKB_HOST_FS": "1", "ZKB_HOST_ASSEMBLY": "1"}, {"ZKB_HOST_ASSEMBLY": "1"}):
        os.environ.update(env)
        try:
            ps = zk.IndependentProofStream()
            top = fri.prove(cuda(cw), ps)
            tree = zk.MerkleTree(cuda(cw), ctx)
            ctx.check(ctx.lib.zkb_merkle_open_ps(tree.h, idx, 9, ps.h))
            tree.close()
            d.append((top, ps.digest()))
        finally:
            for k in env:
                os.environ.pop(k, None)
    assert d[0] == d[1] == d[2] == d[3]


@pytest.mark.parametrize("ncc,ef", [(1, 4), (2, 2), (4, 4), (16, 8), (64, 4), (200, 4)])
def test_device_fs_unusual_parameters(ctx, ncc, ef):
    """num_rounds (fri.rs:40-50) ends at layers from 2^1 to 2^10 values depending on ef / ncc: the tail kernel's small-layer cases."""
    log_n = 12
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    cw = C.coset_lde(w, n, F.GENERATOR, C.synth(77, n // ef))
    fri = zk.FRI(F.GENERATOR, w, n, ef, ncc, ctx)
    ofri = OFRI(F.GENERATOR, w, n, ef, ncc)
    ps, ops = zk.IndependentProofStream(), PS.IndependentProofStream()
    codewords, trees, _ = fastfri.commit(ofri, cw, ops)
    layers = fri.commit(cuda(cw), ps)
    assert len(layers) == ofri.num_rounds()
    assert ps.digest() == ops.digest()
    for r in range(len(layers)):
        assert np.array_equal(layers.codeword(r), codewords[r]), r
    layers.close()


def test_device_fs_back_to_back_calls(ctx):
    """The barrier word, the sequence flag and the device sponge are reused across calls of one context."""
    for k in range(6):
        log_n = (12, 16, 18)[k % 3]
        n, w, cw = _codeword(log_n, 900 + k)
        ps, ops = zk.IndependentProofStream(), PS.IndependentProofStream()
        fastfri.commit(OFRI(F.GENERATOR, w, n, 4, 64), cw, ops)
        zk.FRI(F.GENERATOR, w, n, 4, 64, ctx).commit(cuda(cw), ps).close()
        assert ps.digest() == ops.digest()


@pytest.mark.parametrize("log_n,ncc,B", [(12, 64, 5), (10, 16, 3), (13, 8, 2), (17, 4, 2)])
def test_batch_device_assembly_equals_host_assembly(ctx, log_n, ncc, B):
    """zkb_fri_prove_batch / zkb_merkle_build_batch / zkb_merkle_open_ps_batch: device Fiat-Shamir + objects framed by the kernels
    (k_open_wire, k_leafs_wire) == the round-1 path (host hop per round, host threads frame every node; ZKB_HOST_ASSEMBLY=1) ==
    the single-instance calls, byte for byte, for several tree depths (ragged record alignments)."""
    import torch
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    fri = zk.FRI(F.GENERATOR, w, n, 4, ncc, ctx)
    cws = np.stack([C.coset_lde(w, n, F.GENERATOR, C.synth(4242 + b, n // 4)) for b in range(B)])
    d = cuda(cws.reshape(-1, 2))
    k = 2 * ncc
    rng = np.random.RandomState(log_n)
    open_idx = np.ascontiguousarray(rng.randint(0, n, size=(B, k)).astype(np.uint64))

    def run_batch():
        streams = [zk.SignatureProofStream(b"doc %d" % b) for b in range(B)]
        handles = (ctypes.c_void_p * B)(*[s.h.value for s in streams])
        trees = (ctypes.c_void_p * B)()
        ctx.check(ctx.lib.zkb_merkle_build_batch(ctx.h, d.data_ptr(), n, n, B, trees, handles))
        top = np.empty((B, ncc), dtype=np.uint64)
        ctx.check(ctx.lib.zkb_fri_prove_batch(ctx.h, ctypes.byref(fri.params), d.data_ptr(), n, n, B, handles,
                                              top.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        ctx.check(ctx.lib.zkb_merkle_open_ps_batch(trees, B, open_idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k, handles))
        for i in range(B - 1, -1, -1):
            ctx.lib.zkb_merkle_free(trees[i])
        return top.tolist(), [s.digest() for s in streams]

    dev = run_batch()
    os.environ["ZKB_HOST_ASSEMBLY"] = "1"
    try:
        hst = run_batch()
    finally:
        os.environ.pop("ZKB_HOST_ASSEMBLY", None)
    assert dev[0] == hst[0]
    assert dev[1] == hst[1]
    # and against the single-instance entry points
    for b in range(B):
        ps = zk.SignatureProofStream(b"doc %d" % b)
        tree = zk.MerkleTree(cuda(cws[b]), ctx)
        ps.push((PS.ROOT, tree.root()))
        top = fri.prove(cuda(cws[b]), ps)
        ctx.check(ctx.lib.zkb_merkle_open_ps(tree.h, open_idx[b].ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k, ps.h))
        assert top == dev[0][b]
        assert ps.digest() == dev[1][b], b
        tree.close()


@pytest.mark.parametrize("log_n", [18, 19, 21, 22])
def test_fused_ntt_leaf_pass_parity(ctx, log_n):
    """ZKB_NTT_LEAF_FUSION=1: the LDE's last pass hashes its own output (k_ntt_rr_leaf) - same codeword, same roots, same stream
    bytes as the two-kernel path (final pass widths 6, 7 and 8)."""
    import torch
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    coeffs = cuda(C.synth(0xF00D + log_n, n // 4))
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctx)
    res = []
    for fused in (False, True):
        if fused:
            os.environ["ZKB_NTT_LEAF_FUSION"] = "1"
        try:
            ps = zk.IndependentProofStream()
            layers = fri.lde_commit(coeffs, ps)
            res.append((ps.digest(), [layers.root(r) for r in range(len(layers))], hashlib.sha256(layers.codeword(0).tobytes()).hexdigest()))
            layers.close()
        finally:
            os.environ.pop("ZKB_NTT_LEAF_FUSION", None)
    assert res[0] == res[1]
