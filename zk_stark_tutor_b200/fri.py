"""FRI: the mirror of src/fri.rs (FRI::new / num_rounds / prove, plus the commit and query
phases as separate calls so a caller can keep the layers in HBM)."""
import ctypes

import numpy as np

from . import _lib
from .context import Vec, default_context, from_le16, le16, unpack
from .proof_stream import CODEWORD, LEAFS, PATH, ROOT, PROOF_BYTES
from .field import Field


class FriLayers:
    """zkb_fri_layers: every codeword + pruned Merkle tree of a commit phase, on the device."""

    def __init__(self, ctx, h, keep):
        self.ctx, self.h, self.keep = ctx, h, keep

    def __len__(self):
        return int(self.ctx.lib.zkb_fri_layer_count(self.h))

    def length(self, r):
        return int(self.ctx.lib.zkb_fri_layer_len(self.h, r))

    def root(self, r):
        out = (ctypes.c_uint8 * 64)()
        self.ctx.check(self.ctx.lib.zkb_fri_layer_root(self.h, r, out))
        return bytes(out)

    def codeword(self, r):
        out = np.empty((self.length(r), 2), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zkb_fri_layer_codeword(self.h, r, out.ctypes.data))
        return out

    def query(self, r, indices_c):
        """fri.rs:174-208 payloads: ([(a, b, c)], [(path_a, path_b, path_c)])."""
        k = len(indices_c)
        d_cur = self.length(r).bit_length() - 1
        d_nxt = self.length(r + 1).bit_length() - 1
        idx = (ctypes.c_uint64 * k)(*indices_c)
        leafs = np.empty((k * 3, 2), dtype=np.uint64)
        paths = np.empty(k * (2 * d_cur + d_nxt) * 64, dtype=np.uint8)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        self.ctx.check(self.ctx.lib.zkb_fri_query(self.h, r, idx, k, leafs.ctypes.data_as(u8p), paths.ctypes.data_as(u8p)))
        lv = unpack(leafs)
        raw = paths.tobytes()
        triples, trip_paths, o = [], [], 0
        for s in range(k):
            triples.append((lv[3 * s], lv[3 * s + 1], lv[3 * s + 2]))
            ps = []
            for d in (d_cur, d_cur, d_nxt):
                ps.append([raw[o + 64 * l:o + 64 * (l + 1)] for l in range(d)])
                o += 64 * d
            trip_paths.append(tuple(ps))
        return triples, trip_paths

    def close(self):
        if self.h is not None and self.h.value:
            self.ctx.lib.zkb_fri_layers_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FRI:
    def __init__(self, offset, omega, domain_length, expansion_factor, num_colinearity_tests, ctx=None):
        """FRI::new fri.rs:23-38."""
        self.offset, self.omega = offset, omega
        self.domain_length = domain_length
        self.expansion_factor = expansion_factor
        self.num_colinearity_tests = num_colinearity_tests
        self.ctx = ctx
        p = _lib.FriParams()
        p.offset[:] = list(int(offset).to_bytes(16, "little"))
        p.omega[:] = list(int(omega).to_bytes(16, "little"))
        p.domain_length, p.expansion_factor, p.num_colinearity_tests = domain_length, expansion_factor, num_colinearity_tests
        self.params = p

    def num_rounds(self):
        """fri.rs:40-50"""
        return int(_lib.lib().zkb_fri_num_rounds(ctypes.byref(self.params)))

    @staticmethod
    def sample_indices(seed, size, reduced_size, number):
        """fri.rs:85-113"""
        assert number <= reduced_size, "Cannot sample more indices than available in the last codeword"
        out = (ctypes.c_uint64 * number)()
        rc = _lib.lib().zkb_fri_sample_indices((ctypes.c_uint8 * len(seed)).from_buffer_copy(seed), len(seed), size, reduced_size, number, out)
        assert rc == 0
        return list(out)

    def fold(self, codeword, alpha, offset=None, omega=None):
        """the split-and-fold step fri.rs:150-159"""
        ctx = self.ctx or default_context()
        v = Vec(codeword)
        out, optr = ctx.out_like(v, v.n // 2)
        ctx.check(ctx.lib.zkb_fri_fold(ctx.h, v.ptr, v.n, le16(alpha), le16(self.offset if offset is None else offset),
                                       le16(self.omega if omega is None else omega), optr))
        return ctx.finish(v, out)

    def lde_commit(self, coefficients, proof_stream):
        """stark.rs:500-522 in one call: fast_coset_evaluate(omega, domain_length, offset, poly)
        then FRI::commit, the codeword never leaving HBM (layer 0 of the result)."""
        return self.commit(coefficients, proof_stream, _from_coefficients=True)

    def commit(self, codeword, proof_stream, _from_coefficients=False):
        """FRI::commit fri.rs:115-172 against any object with push() / fiat_shamir_prover():
        the Fiat-Shamir hop is a host callback, so custom proof streams keep working."""
        ctx = self.ctx or default_context()
        v = Vec(codeword)
        from .proof_stream import IndependentProofStream
        if isinstance(proof_stream, IndependentProofStream) and not getattr(proof_stream, "force_python", False):
            h = ctypes.c_void_p()
            fn = ctx.lib.zkb_lde_fri_commit_ps if _from_coefficients else ctx.lib.zkb_fri_commit_ps
            ctx.check(fn(ctx.h, ctypes.byref(self.params), v.ptr if v.n else None, v.n, proof_stream.h, ctypes.byref(h)))
            proof_stream.objects = None
            return FriLayers(ctx, h, v)
        field = Field()
        err = []

        def cb(_user, _round, root_p, want_alpha, alpha_out):
            try:
                proof_stream.push((ROOT, bytes(root_p[:64])))                       # fri.rs:136-137
                if want_alpha:
                    alpha = field.sample(proof_stream.fiat_shamir_prover(PROOF_BYTES))   # fri.rs:145-146
                    for i, b in enumerate(alpha.to_bytes(16, "little")):
                        alpha_out[i] = b
                return 0
            except Exception as e:        # noqa: BLE001 - reported through the C status code
                err.append(e)
                return 1

        h = ctypes.c_void_p()
        fn = ctx.lib.zkb_lde_fri_commit if _from_coefficients else ctx.lib.zkb_fri_commit
        rc = fn(ctx.h, ctypes.byref(self.params), v.ptr if v.n else None, v.n, _lib.FS_CALLBACK(cb), None, ctypes.byref(h))
        if err:
            raise err[0]
        ctx.check(rc)
        layers = FriLayers(ctx, h, v)
        proof_stream.push((CODEWORD, unpack(layers.codeword(len(layers) - 1))))     # fri.rs:166
        return layers

    def prove(self, codeword, proof_stream):
        """FRI::prove fri.rs:210-248 -> top-level indices.  With the library's own proof
        stream the whole protocol runs in one C call (zkb_fri_prove)."""
        ctx = self.ctx or default_context()
        v = Vec(codeword)
        if v.n != self.domain_length:
            raise AssertionError("Length of the domain doesnt match the length of initial codeword")
        from .proof_stream import IndependentProofStream
        if isinstance(proof_stream, IndependentProofStream) and not getattr(proof_stream, "force_python", False):
            top = (ctypes.c_uint64 * self.num_colinearity_tests)()
            ctx.check(ctx.lib.zkb_fri_prove(ctx.h, ctypes.byref(self.params), v.ptr, v.n, proof_stream.h, top))
            proof_stream.objects = None          # the transcript lives in the C stream; read it with digest()
            return list(top)
        layers = self.commit(v.keep, proof_stream)
        try:
            R = len(layers)
            top = self.sample_indices(proof_stream.fiat_shamir_prover(PROOF_BYTES), layers.length(1), layers.length(R - 1),
                                      self.num_colinearity_tests)
            indices = list(top)
            for r in range(R - 1):
                indices = [i % (layers.length(r) // 2) for i in indices]            # fri.rs:234-237
                triples, paths = layers.query(r, indices)
                for t in triples:
                    proof_stream.push((LEAFS, t))
                for pa, pb, pc in paths:
                    proof_stream.push((PATH, pa))
                    proof_stream.push((PATH, pb))
                    proof_stream.push((PATH, pc))
            return top
        finally:
            layers.close()
