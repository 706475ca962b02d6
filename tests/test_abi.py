"""The C-ABI library loads and exports exactly the symbols include/zkb200.h declares
(no compute calls: there is no GPU here), and the host-only entry points (field scalars,
hashes, proof stream, sample_indices, Merkle verify) agree with the oracle / the
reference's known-answer tests."""
import ctypes
import os
import re
import subprocess

import pytest

from oracle import field as F, fri as ofri, merkle as omerkle, proof_stream as ops
from zk_stark_tutor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def _buf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr)) - {"zkb_fs_callback"}
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T zkb_" in ln}
    assert declared == exported, (declared ^ exported)
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))


def test_no_device_means_no_context(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    assert L.zkb_ctx_create(0, None, ctypes.byref(ctx)) == -1      # ZKB_ERR_CUDA: no CPU fallback
    assert not ctx.value


def test_field_scalars(L):
    out = (ctypes.c_uint8 * 16)()
    # field.rs:193-200
    assert L.zkb_primitive_nth_root(256, out) == 0
    assert F.from_le16(bytes(out)) == 178902808384765167578311106676137348214
    assert L.zkb_primitive_nth_root(2, out) == 0
    assert F.from_le16(bytes(out)) == F.P - 1
    assert L.zkb_primitive_nth_root(3, out) != 0
    for n in (1, 4, 1 << 16, 1 << 24, 1 << 40):
        L.zkb_primitive_nth_root(n, out)
        assert F.from_le16(bytes(out)) == F.primitive_nth_root(n)
    L.zkb_field_generator(out)
    assert F.from_le16(bytes(out)) == 85408008396924667383611388730472331217
    # field_element.rs:153-156, :205-208, :292
    a, b = 49789714223038013592473676705012096123, 6534789852937546098347957826345234
    L.zkb_field_mul(_buf(F.to_le16(a)), _buf(F.to_le16(b)), out)
    assert F.from_le16(bytes(out)) == 105250150227149389100670877502232671566
    L.zkb_field_inv(_buf(F.to_le16(256)), out)
    assert F.from_le16(bytes(out)) == 269441264731518542713518780764053831681
    L.zkb_field_pow(_buf(F.to_le16(6534789852937546098)), 501209126122, out)
    assert F.from_le16(bytes(out)) == 256557788041265930815463337858691703671
    # field.rs:230-240
    for data, want in ((bytes.fromhex("6c9c4992"), 1822181778), (bytes.fromhex("ac4cd3be"), 2890716094)):
        L.zkb_field_sample(_buf(data), len(data), out)
        assert F.from_le16(bytes(out)) == want
    for data in (b"\xff" * 16, b"\x01" + b"\xff" * 16, bytes(range(40)), b"\xcb\x80" + b"\x00" * 13 + b"\x01", b"\xcb\x80" + b"\x00" * 13 + b"\x02"):
        L.zkb_field_sample(_buf(data), len(data), out)
        assert F.from_le16(bytes(out)) == F.sample(data)


def test_hashes(L):
    out = (ctypes.c_uint8 * 64)()
    # blake2b512.rs:22-30
    L.zkb_blake2b512(_buf(b"\x00"), 1, out)
    assert bytes(out).hex().startswith("2fa3f686df876995167e7c2e5d74c4c7")
    for n in (0, 1, 64, 127, 128, 129, 255, 256, 257, 1000):
        msg = bytes((i * 7 + 3) & 0xFF for i in range(n))
        L.zkb_blake2b512(_buf(msg) if n else None, n, out)
        assert bytes(out) == omerkle.blake2b512(msg)
    for n in (0, 1, 135, 136, 137, 272, 1000):
        msg = bytes((i * 5 + 1) & 0xFF for i in range(n))
        for olen in (4, 32, 64, 136, 137, 300):
            o = (ctypes.c_uint8 * olen)()
            L.zkb_shake256(_buf(msg) if n else None, n, o, olen)
            assert bytes(o) == ops.shake256(msg, olen)


def test_merkle_verify_host(L):
    vals = [5462, 456, 652, 23409]          # merkle_root.rs:160-244
    root = omerkle.commit(vals)
    for i in range(4):
        path = omerkle.open_(i, vals)
        pb = b"".join(path)
        assert L.zkb_merkle_verify(_buf(root), i, _buf(pb), len(path), _buf(F.to_le16(vals[i]))) == 1
        assert L.zkb_merkle_verify(_buf(root), i ^ 1, _buf(pb), len(path), _buf(F.to_le16(vals[i]))) == 0
        assert L.zkb_merkle_verify(_buf(root), i, _buf(pb), len(path), _buf(F.to_le16(vals[i] + 1))) == 0


def test_sample_indices_kat(L):
    seed = bytes.fromhex("d4b6e8af1114859c1c24b6496a3aef2f55a21105bc103af7e12dc3b2c101fe66")   # fri.rs:438-447
    out = (ctypes.c_uint64 * 17)()
    assert L.zkb_fri_sample_indices(_buf(seed), len(seed), 128, 128, 17, out) == 0
    assert list(out) == [40, 121, 5, 113, 97, 68, 126, 88, 26, 82, 81, 91, 93, 125, 10, 57, 48]
    out = (ctypes.c_uint64 * 64)()
    assert L.zkb_fri_sample_indices(_buf(seed), len(seed), 1 << 23, 512, 64, out) == 0
    assert list(out) == ofri.FRI.sample_indices(seed, 1 << 23, 512, 64)


def test_num_rounds(L):
    for n, ef, ncc in ((256, 4, 17), (4096, 4, 64), (1 << 24, 4, 64), (1 << 16, 4, 64), (8, 4, 1), (4, 4, 1)):
        p = _lib.FriParams()
        p.domain_length, p.expansion_factor, p.num_colinearity_tests = n, ef, ncc
        assert L.zkb_fri_num_rounds(ctypes.byref(p)) == ofri.FRI(F.GENERATOR, 1, n, ef, ncc).num_rounds()


def test_proof_stream_bytes(L):
    for doc in (None, b"a document"):
        ps = ctypes.c_void_p()
        assert L.zkb_ps_create(_buf(doc) if doc else None, len(doc) if doc else 0, 1 if doc else 0, ctypes.byref(ps)) == 0
        ref = ops.SignatureProofStream(doc) if doc else ops.IndependentProofStream()
        ch = (ctypes.c_uint8 * 32)()
        L.zkb_ps_fiat_shamir(ps, 32, ch)
        assert bytes(ch) == ref.fiat_shamir_prover(32)
        root = bytes(range(64))
        L.zkb_ps_push_root(ps, _buf(root), 64); ref.push((ops.ROOT, root))
        L.zkb_ps_fiat_shamir(ps, 32, ch)
        assert bytes(ch) == ref.fiat_shamir_prover(32)       # zero header while only Roots
        cw = [3, F.P - 1, 1 << 100]
        L.zkb_ps_push_codeword(ps, _buf(b"".join(F.to_le16(v) for v in cw)), 3); ref.push((ops.CODEWORD, cw))
        path = [bytes([i]) * 64 for i in range(5)]
        L.zkb_ps_push_path(ps, _buf(b"".join(path)), 5); ref.push((ops.PATH, path))
        L.zkb_ps_push_leafs(ps, _buf(F.to_le16(1)), _buf(F.to_le16(2)), _buf(F.to_le16(F.P - 2))); ref.push((ops.LEAFS, (1, 2, F.P - 2)))
        L.zkb_ps_push_value(ps, _buf(F.to_le16(77))); ref.push((ops.VALUE, 77))
        want = ref.digest()
        n = L.zkb_ps_digest(ps, None, 0)
        assert n == len(want)
        buf = (ctypes.c_uint8 * n)()
        L.zkb_ps_digest(ps, buf, n)
        assert bytes(buf) == want
        L.zkb_ps_fiat_shamir(ps, 32, ch)
        assert bytes(ch) == ref.fiat_shamir_prover(32)
        L.zkb_ps_free(ps)


def test_proof_stream_incremental_sponge(L):
    """zkb_ps_fiat_shamir keeps a running SHAKE256 sponge over the transcript (absorbs only the bytes pushed since the last
    challenge; restarts once when the header flips to the field order): a challenge after EVERY push of a long random object
    sequence equals SHAKE256 of the whole stream, as proof_stream.rs:36-40 computes it."""
    import random
    r = random.Random(5)
    for doc in (None, b"doc"):
        ps = ctypes.c_void_p()
        assert L.zkb_ps_create(_buf(doc) if doc else None, len(doc) if doc else 0, 1 if doc else 0, ctypes.byref(ps)) == 0
        ref = ops.SignatureProofStream(doc) if doc else ops.IndependentProofStream()
        ch = (ctypes.c_uint8 * 32)()
        for step in range(120):
            kind = r.choice([0, 0, 0, 2, 2] if step < 40 else [0, 1, 2, 3, 4])        # Roots / Paths only first: zero header
            if kind == 0:
                root = bytes(r.randrange(256) for _ in range(64))
                L.zkb_ps_push_root(ps, _buf(root), 64); ref.push((ops.ROOT, root))
            elif kind == 1:
                cw = [r.randrange(F.P) for _ in range(r.randrange(0, 40))]
                L.zkb_ps_push_codeword(ps, _buf(b"".join(F.to_le16(v) for v in cw)) if cw else None, len(cw)); ref.push((ops.CODEWORD, cw))
            elif kind == 2:
                path = [bytes(r.randrange(256) for _ in range(64)) for _ in range(r.randrange(1, 13))]
                L.zkb_ps_push_path(ps, _buf(b"".join(path)), len(path)); ref.push((ops.PATH, path))
            elif kind == 3:
                t = tuple(r.randrange(F.P) for _ in range(3))
                L.zkb_ps_push_leafs(ps, _buf(F.to_le16(t[0])), _buf(F.to_le16(t[1])), _buf(F.to_le16(t[2]))); ref.push((ops.LEAFS, t))
            else:
                v = r.randrange(F.P)
                L.zkb_ps_push_value(ps, _buf(F.to_le16(v))); ref.push((ops.VALUE, v))
            if step % 3 != 2:                                   # some pushes without a challenge in between
                L.zkb_ps_fiat_shamir(ps, 32, ch)
                assert bytes(ch) == ref.fiat_shamir_prover(32), step
        big = (ctypes.c_uint8 * 300)()
        L.zkb_ps_fiat_shamir(ps, 300, big)
        assert bytes(big) == ref.fiat_shamir_prover(300)
        L.zkb_ps_free(ps)


def test_proof_stream_raw_objects(L):
    """zkb_ps_push_object: objects the caller serialised itself (any path node size, stark.rs:785-808's test shape) give the oracle's
    wire bytes and header rule (order only once a field-carrying object is present, proof_stream_enum.rs:161-190)."""
    ps = ctypes.c_void_p()
    assert L.zkb_ps_create(None, 0, 0, ctypes.byref(ps)) == 0
    ref = ops.IndependentProofStream()

    def both(code, payload, obj):
        assert L.zkb_ps_push_object(ps, code, _buf(payload) if payload else None, len(payload)) == 0
        ref.push(obj)
        n = L.zkb_ps_digest(ps, None, 0)
        buf = (ctypes.c_uint8 * n)()
        L.zkb_ps_digest(ps, buf, n)
        assert bytes(buf) == ref.digest()
    both(0, bytes([0x49, 0x6e, 0x20, 0x74]), (ops.ROOT, bytes([0x49, 0x6e, 0x20, 0x74])))
    nodes = [bytes([0x49, 0x6e, 0x20, 0x74]), bytes([0x01, 0x6b, 0xfe, 0x25, 0x99])]
    both(2, b"".join(len(x).to_bytes(8, "big") + x for x in nodes), (ops.PATH, nodes))
    both(1, b"", (ops.CODEWORD, []))                                   # an empty codeword carries no field: header still zero
    both(4, (2).to_bytes(16, "big"), (ops.VALUE, 2))                   # now the header is the field order
    both(3, b"".join(v.to_bytes(16, "big") for v in (1, 5, 10)), (ops.LEAFS, (1, 5, 10)))
    both(1, b"".join(v.to_bytes(16, "big") for v in (20, 100)), (ops.CODEWORD, [20, 100]))
    assert L.zkb_ps_push_object(ps, 5, _buf(b"x"), 1) == -2             # "Unknown code" (proof_stream_enum.rs:62)
    L.zkb_ps_free(ps)


def test_context_calls_fail_without_a_context(L):
    """no CUDA device in the CPU container: zkb_ctx_create fails (there is no CPU fallback) and context-taking calls reject NULL"""
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert L.zkb_ctx_create(0, None, ctypes.byref(h)) != 0 and not h.value
    assert L.zkb_ctx_assembly_threads(None, 4) == -2
    assert L.zkb_air_create(None, None, ctypes.byref(h)) == -2
    assert L.zkb_trace_lde_batch(None, None, 0, 0, None, 0, None, None, 0, 0, None, 0, None) == -2
    assert L.zkb_coset_degree_batch(None, None, None, 0, 0, 0, None) == -2
    assert L.zkb_air_combination(None, None, None, 0, None, None, None) == -2
    assert L.zkb_stark_prove_batch(None, None, None, 0, None, None, None, None, None) == -2


def test_stark_shape_binding_matches_the_header():
    """the ctypes mirror of zkb_stark_shape (zkb_stark_prove_batch) has the C struct's layout: sizes and offsets from a compiled probe"""
    import shutil
    import subprocess
    import tempfile
    from zk_stark_tutor_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = ('#include <stddef.h>\n#include <stdio.h>\n#include "zkb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(zkb_stark_shape), '
           'offsetof(zkb_stark_shape, rnd_poly_len), offsetof(zkb_stark_shape, tq_degree_bounds), offsetof(zkb_stark_shape, lagrange), '
           'offsetof(zkb_stark_shape, fri), offsetof(zkb_stark_shape, proof_bytes));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    S = _lib.StarkShape
    assert got == [ctypes.sizeof(S), S.rnd_poly_len.offset, S.tq_degree_bounds.offset, S.lagrange.offset, S.fri.offset, S.proof_bytes.offset]
