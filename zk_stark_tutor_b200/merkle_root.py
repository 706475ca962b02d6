"""MerkleRoot: the mirror of src/merkle_root.rs for T = FieldElement (the only T the crate
uses on this path).  commit / open / verify keep the reference's static-method shape;
MerkleTree is the retained-tree handle that makes batched openings O(log n) instead of the
reference's full rebuild per MerkleRoot::open (merkle_root.rs:55-66)."""
import ctypes

import numpy as np

from .context import Vec, default_context, le16


class MerkleTree:
    """zkb_merkle_build: keeps the (pruned) tree on the device."""

    def __init__(self, leafs, ctx=None):
        self.ctx = ctx or default_context()
        self.v = Vec(leafs)                      # keeps a device tensor alive while the tree references it
        self.n = self.v.n
        h = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.zkb_merkle_build(self.ctx.h, self.v.ptr if self.n else None, self.n, ctypes.byref(h)))
        self.h = h

    def root(self):
        out = (ctypes.c_uint8 * 64)()
        self.ctx.lib.zkb_merkle_root(self.h, out)
        return bytes(out)

    def open_many(self, indices):
        """[path(i) for i in indices]; path = list of 64-byte nodes, leaf sibling first."""
        k = len(indices)
        depth = self.n.bit_length() - 1
        idx = (ctypes.c_uint64 * max(k, 1))(*indices)
        out = np.empty(max(k * depth * 64, 1), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.zkb_merkle_open(self.h, idx, k, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        raw = out.tobytes()
        return [[raw[(s * depth + l) * 64:(s * depth + l + 1) * 64] for l in range(depth)] for s in range(k)]

    def open(self, index):
        return self.open_many([index])[0]

    def open_into(self, indices, proof_stream):
        """The opening loop of Stark::prove (stark.rs:546-560): for i in indices push Value(leaf i)
        then Path(open(i)) onto the library's proof stream, one batched opening."""
        k = len(indices)
        idx = (ctypes.c_uint64 * max(k, 1))(*indices)
        self.ctx.check(self.ctx.lib.zkb_merkle_open_ps(self.h, idx, k, proof_stream.h))
        proof_stream.objects = None

    def close(self):
        if self.h is not None and self.h.value:
            self.ctx.lib.zkb_merkle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MerkleRoot:
    @staticmethod
    def commit(leafs, ctx=None):
        """MerkleRoot::commit merkle_root.rs:21-32 -> 64-byte root."""
        ctx = ctx or default_context()
        v = Vec(leafs)
        out = (ctypes.c_uint8 * 64)()
        ctx.check(ctx.lib.zkb_merkle_commit(ctx.h, v.ptr if v.n else ctypes.c_void_p(8), v.n, out))
        return bytes(out)

    @staticmethod
    def open(index, leafs, ctx=None):
        """MerkleRoot::open merkle_root.rs:55-66."""
        t = MerkleTree(leafs, ctx)
        try:
            return t.open(index)
        finally:
            t.close()

    @staticmethod
    def verify(root, index, path, leaf):
        """MerkleRoot::verify merkle_root.rs:89-95 (host)."""
        from . import _lib
        pb = b"".join(path)
        return bool(_lib.lib().zkb_merkle_verify((ctypes.c_uint8 * 64).from_buffer_copy(root), index,
                                                 (ctypes.c_uint8 * max(len(pb), 1)).from_buffer_copy(pb or b"\0"),
                                                 len(path), le16(leaf)))
