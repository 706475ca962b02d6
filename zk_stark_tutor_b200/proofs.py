"""Batches of RPSSS-shaped proofs (BASELINE configs[4]): the hot-path call
sequence of one Stark::prove (stark.rs:363-562) at the tutorial's signature parameters -
every committed polynomial is LDE'd to the 4096-point FRI domain and Merkle-committed
(stark.rs:373-381, 431-436), the combination polynomial is LDE'd and goes through FRI::prove
(stark.rs:520-531), and every committed codeword is opened at the quadrupled, sorted FRI indices
(stark.rs:534-560) - executed as a handful of C-ABI calls per proof, with the codewords, trees
and FRI layers never leaving HBM.

What is NOT here (out of scope): the AIR itself (MPolynomial constraint evaluation,
quotient construction) - the caller hands in the coefficient vectors it would have produced
(boundary quotients, randomizer, weighted combination).  The transcript layout, Fiat-Shamir
hops and the proof bytes are exactly those of the reference for these inputs.

Independent proofs are independent units: `partition` deals them round-robin to ranks (no data
crosses NVLink) and `ProofPipeline` keeps several in flight per GPU, because at a 4096-point
domain every kernel is latency-bound and one proof alone leaves most of the GPU idle."""
import ctypes
import threading

import numpy as np

from .columns import partition  # noqa: F401  (same round-robin deal as trace columns)
from .context import le16


class ProofShape:
    """Sizes of one RPSSS signature proof (rpsss.rs:24-40, stark.rs:186-200)."""

    def __init__(self, omicron_len=1024, expansion_factor=4, num_colinearity_tests=64, n_registers=2,
                 trace_len=284, tc_degree=3):
        self.omicron_len = omicron_len
        self.ef = expansion_factor
        self.ncc = num_colinearity_tests
        self.n_registers = n_registers
        self.fri_len = omicron_len * expansion_factor
        self.bq_len = trace_len - 2                                              # boundary-quotient coefficients
        self.max_degree = (1 << (tc_degree * (trace_len - 1)).bit_length()) - 1  # stark.rs:186-200
        self.comb_len = self.max_degree + 1

    def column_lengths(self):
        """coefficient counts of the committed polynomials: the registers' boundary quotients, then the randomizer"""
        return [self.bq_len] * self.n_registers + [self.max_degree + 1]


def quadrupled_indices(top, fri_len, ef):
    """stark.rs:534-542: duplicated (+ef), quadrupled (+fri_len/2), sorted, NOT deduplicated."""
    dup = list(top) + [(i + ef) % fri_len for i in top]
    return sorted(dup + [(i + fri_len // 2) % fri_len for i in dup])


def prove_hot_path(ctx, fri, shape, columns, combination, proof_stream):
    """One proof.  columns: the committed coefficient vectors ((k, 2) uint64 arrays / tensors, host or
    device); combination: the coefficients of the weighted combination; proof_stream: a
    zk.IndependentProofStream / SignatureProofStream.  Returns the top-level FRI indices."""
    import torch
    lib = ctx.lib
    n = shape.fri_len
    omega, offset = le16(fri.omega), le16(fri.offset)
    dev = torch.device("cuda", ctx.device)
    # one device buffer for all codewords of this proof: the columns, then the combination
    cws = torch.empty((len(columns) + 1, n, 2), dtype=torch.int64, device=dev)
    trees = []
    try:
        for k, col in enumerate(columns):                                   # stark.rs:373-381, 431-436
            v = _vec(col)
            ctx.check(lib.zkb_coset_lde(ctx.h, omega, n, offset, v[0], v[1], cws[k].data_ptr()))
            h = ctypes.c_void_p()
            ctx.check(lib.zkb_merkle_build(ctx.h, cws[k].data_ptr(), n, ctypes.byref(h)))
            trees.append(h)
            root = (ctypes.c_uint8 * 64)()
            lib.zkb_merkle_root(h, root)
            lib.zkb_ps_push_root(proof_stream.h, root, 64)
        v = _vec(combination)                                               # stark.rs:520-531
        ctx.check(lib.zkb_coset_lde(ctx.h, omega, n, offset, v[0], v[1], cws[len(columns)].data_ptr()))
        top = (ctypes.c_uint64 * shape.ncc)()
        ctx.check(lib.zkb_fri_prove(ctx.h, ctypes.byref(fri.params), cws[len(columns)].data_ptr(), n, proof_stream.h, top))
        quad = quadrupled_indices(list(top), n, shape.ef)                   # stark.rs:534-542
        idx = (ctypes.c_uint64 * len(quad))(*quad)
        for h in trees:                                                     # stark.rs:546-560
            ctx.check(lib.zkb_merkle_open_ps(h, idx, len(quad), proof_stream.h))
        proof_stream.objects = None
        return list(top)
    finally:
        try:                                   # torch's allocator does not know the context's own stream: drain it before the tensors go
            ctx.sync()
        except Exception:                      # noqa: BLE001 - the original error is the one to report
            pass
        for h in trees:
            lib.zkb_merkle_free(h)


def pack_batch(shape, proofs):
    """The coefficient vectors of a batch in the layout prove_hot_path_batch takes: a (K + 1, B, comb_len, 2)
    uint64 array, zero padded - plane t < K holds committed polynomial t of every proof, plane K the combinations.
    (Zero padding does not change an LDE: the extra coefficients are zero.)"""
    K, B = len(shape.column_lengths()), len(proofs)
    out = np.zeros((K + 1, B, shape.comb_len, 2), dtype=np.uint64)
    for b, (cols, comb) in enumerate(proofs):
        for t, col in enumerate(cols):
            a = np.ascontiguousarray(col).view(np.uint64).reshape(-1, 2)
            out[t, b, :a.shape[0]] = a
        a = np.ascontiguousarray(comb).view(np.uint64).reshape(-1, 2)
        out[K, b, :a.shape[0]] = a
    return out


def prove_hot_path_batch(ctx, fri, shape, packed, proof_streams):
    """prove_hot_path for B proofs in lockstep (csrc/batch.cu): every launch carries all B instances.
    packed: pack_batch(...) as a numpy array / pinned or CUDA tensor of shape (K + 1, B, comb_len, 2);
    proof_streams: B library proof streams.  Returns the B lists of top-level FRI indices.  The bytes each
    stream ends up with are identical to prove_hot_path's."""
    import torch
    lib = ctx.lib
    n, K, B, nc = shape.fri_len, len(shape.column_lengths()), len(proof_streams), shape.comb_len
    ptr = packed.data_ptr() if hasattr(packed, "data_ptr") else np.ascontiguousarray(packed).ctypes.data
    assert tuple(packed.shape) == (K + 1, B, nc, 2)
    dev = torch.device("cuda", ctx.device)
    cws = torch.empty(((K + 1) * B, n, 2), dtype=torch.int64, device=dev)       # plane-major like `packed`
    omega, offset = le16(fri.omega), le16(fri.offset)
    ctx.check(lib.zkb_coset_lde_batch(ctx.h, omega, n, offset, ptr, nc, nc, cws.data_ptr(), n, (K + 1) * B))   # stark.rs:373-381, 431-436, 520-522
    trees = (ctypes.c_void_p * (K * B))()
    handles = [p.h.value for p in proof_streams]
    ps_arr = (ctypes.c_void_p * B)(*handles)
    ps_of_tree = (ctypes.c_void_p * (K * B))(*(handles * K))                    # tree t*B + b belongs to proof b
    ctx.check(lib.zkb_merkle_build_batch(ctx.h, cws.data_ptr(), n, n, K * B, trees, ps_of_tree))   # commits + Root pushes, tree order
    try:
        top = np.empty((B, shape.ncc), dtype=np.uint64)
        ctx.check(lib.zkb_fri_prove_batch(ctx.h, ctypes.byref(fri.params), cws[K * B].data_ptr(), n, n, B, ps_arr,
                                          top.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))               # stark.rs:520-531
        nn, ef = np.uint64(n), np.uint64(shape.ef)
        dup = np.concatenate([top, (top + ef) % nn], axis=1)                     # stark.rs:534-542, all proofs at once
        quad = np.sort(np.concatenate([dup, (dup + nn // np.uint64(2)) % nn], axis=1), axis=1)
        k = quad.shape[1]
        idx = np.ascontiguousarray(np.broadcast_to(quad[None, :, :], (K, B, k)))
        ctx.check(lib.zkb_merkle_open_ps_batch(trees, K * B, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k, ps_of_tree))   # stark.rs:546-560
        for p in proof_streams:
            p.objects = None
        return top.tolist()
    finally:
        try:                                   # torch's allocator does not know the context's own stream: drain it before the tensors go
            ctx.sync()
        except Exception:                      # noqa: BLE001 - the original error is the one to report
            pass
        for i in range(K * B - 1, -1, -1):           # tree 0 owns the shared arena: free it last
            if trees[i]:
                lib.zkb_merkle_free(trees[i])


def _vec(x):
    """(pointer, n) of a (n, 2) 64-bit numpy array or torch tensor (host or device)."""
    if hasattr(x, "data_ptr"):
        return x.data_ptr(), x.shape[0]
    a = np.ascontiguousarray(x).view(np.uint64).reshape(-1, 2)
    return a.ctypes.data, a.shape[0]


class ProofPipeline:
    """`lanes` proofs in flight on ONE GPU: one context (own CUDA stream) and one host thread per
    lane; the C calls release the GIL."""

    def __init__(self, device, shape, offset, omega, lanes=8, assembly_threads=None):
        """assembly_threads: host threads per batched call for proof-stream assembly (zkb_ctx_assembly_threads; None = the library's 16).
        With several lanes the lanes themselves are the host parallelism, so a few threads per call do better on a 16-core host."""
        import zk_stark_tutor_b200 as zk
        self.zk, self.shape = zk, shape
        self.ctxs = [zk.Context(device, stream="own") for _ in range(lanes)]
        if lanes > 1:
            for c in self.ctxs:      # more contexts in flight than idle cores: sleep while the GPU works instead of spinning on a core each
                c.check(c.lib.zkb_ctx_blocking_sync(c.h, 1))
        if assembly_threads:
            for c in self.ctxs:
                c.check(c.lib.zkb_ctx_assembly_threads(c.h, int(assembly_threads)))
        self.fris = [zk.FRI(offset, omega, shape.fri_len, shape.ef, shape.ncc, c) for c in self.ctxs]

    def run(self, proofs, make_stream, keep_digest=False):
        """proofs: list of (columns, combination).  Returns per proof (proof bytes length, digest or None)."""
        results = [None] * len(proofs)
        errors = []

        def worker(lane):
            try:
                ctx, fri = self.ctxs[lane], self.fris[lane]
                for i in range(lane, len(proofs), len(self.ctxs)):
                    ps = make_stream()
                    prove_hot_path(ctx, fri, self.shape, proofs[i][0], proofs[i][1], ps)
                    results[i] = (int(ctx.lib.zkb_ps_digest(ps.h, None, 0)), ps.digest() if keep_digest else None)
                    ps.close()
            except Exception as e:       # noqa: BLE001 - re-raised in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(self.ctxs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for c in self.ctxs:
            c.sync()
        if errors:
            raise errors[0]
        return results

    def run_batched(self, batches, make_stream, keep_digest=False):
        """batches: list of pack_batch(...) arrays / tensors, each a group of proofs that advances in lockstep
        (prove_hot_path_batch); the lanes take batches round-robin, so one lane's host-side proof assembly
        overlaps another lane's kernels.  Returns per proof (proof bytes length, digest or None), batch-major."""
        results = [None] * len(batches)
        errors = []

        def worker(lane):
            try:
                ctx, fri = self.ctxs[lane], self.fris[lane]
                for i in range(lane, len(batches), len(self.ctxs)):
                    streams = [make_stream() for _ in range(batches[i].shape[1])]
                    prove_hot_path_batch(ctx, fri, self.shape, batches[i], streams)
                    results[i] = [(int(ctx.lib.zkb_ps_digest(ps.h, None, 0)), ps.digest() if keep_digest else None) for ps in streams]
                    for ps in streams:
                        ps.close()
            except Exception as e:       # noqa: BLE001 - re-raised in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(self.ctxs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for c in self.ctxs:
            c.sync()
        if errors:
            raise errors[0]
        return [r for batch in results for r in batch]

    def close(self):
        for c in self.ctxs:
            c.close()
