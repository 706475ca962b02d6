"""Stark::prove (src/stark/stark.rs:276-563) with the hot path AND its evaluation-form middle on the B200:

    trace interpolation            fast_interpolate_domain            (stark.rs:303-326)   GPU NTT products
    boundary quotients             fast_coset_divide                  (stark.rs:331-360)   GPU
    LDE + Merkle commit            fast_coset_evaluate, commit        (stark.rs:366-386, 424-445)   GPU, codewords + trees stay in HBM
    transition quotients, x^shift products, weighted combination, LDE of the combination
                                   (stark.rs:388-519)                 ONE pointwise kernel over the committed codewords
                                                                      (zkb_air_combination, csrc/air.cu) instead of symbolic
                                                                      polynomial arithmetic + another LDE
    FRI::prove                     (stark.rs:522)                     GPU, from the device codeword
    openings                       (stark.rs:546-560)                 one batched opening per committed tree

The transcript, Fiat-Shamir hops and proof bytes are those of the reference for the same randomness: the
evaluation-form route computes the same polynomials' values (exact field arithmetic).  The degree bookkeeping
(stark.rs:115-258) is integer logic on the host, restated here because the prover needs the shifts.
Rescue-Prime / RPSSS themselves (the AIR's author) are out of scope: the caller hands in the trace,
the constraint dictionaries and the boundary conditions.  The reference draws randomizers from thread_rng
(stark.rs:283, 428); here `rng(n) -> n bytes` is a parameter so proofs are reproducible."""
import numpy as np

from . import fft
from .air import air_combination
from .context import P, default_context, pack, unpack
from .field import Field
from .fri import FRI
from .merkle_root import MerkleTree
from .proof_stream import PROOF_BYTES, ROOT


class PrefixInterpolator:
    """The polynomial of degree < L through (root^i, values[i]), i < L, where the points are a PREFIX of the order-N
    subgroup <root> - the shape of Stark's trace domain (stark.rs:305-326: omicron^0 .. omicron^(trace_len-1)).

    The reference interpolates with a divide-and-conquer over arbitrary points (fast_interpolate_domain,
    ntt_arithmetics.rs:172-237: O(L log^2 L) products plus schoolbook remainders).  On a subgroup prefix the same - unique -
    polynomial is the remainder of ANY degree < N polynomial that takes the values on the prefix, modulo the prefix's
    zerofier Z:   p = iNTT(values || 0...) mod Z.   Z and the power-series inverse of its reversal depend only on (L, N) and
    are computed once; per column that leaves one iNTT and two NTT products (fast division), all on the GPU.
    `mul(a, b)`, `intt(values)` are injected so the algorithm can be checked on the CPU against the oracle."""

    def __init__(self, L, N, zerofier, mul, intt):
        assert 0 < L <= N
        self.L, self.N, self.array_mode = L, N, False
        self.mul, self.intt = mul, intt
        self.m = N - L                                    # number of quotient coefficients for a dividend of degree < N
        if self.m:                                        # (L == N: the prefix is the whole subgroup, the iNTT is the answer)
            assert len(zerofier) >= L + 1 and zerofier[L] == 1 and not any(zerofier[L + 1:]), "Z must be monic of degree L"
            self.Z = list(zerofier[:L + 1])
            f = self.Z[::-1][:self.m]                      # rev(Z) mod x^m; f[0] = 1
            f += [0] * (self.m - len(f))
            g = [1]
            while len(g) < self.m:                         # Newton: g <- g * (2 - f*g) mod x^(2 len g)
                k = min(2 * len(g), self.m)
                t = [(-c) % P for c in self._mul_trunc(f[:k], g, k)]
                t[0] = (t[0] + 2) % P
                g = self._mul_trunc(g, t, k)
            self.inv_rev_Z = g
            self.Z_arr, self.inv_arr = pack(self.Z), pack(g)

    def use_arrays(self, mul, intt):
        """switch to (n, 2) uint64 arrays end to end (mul / intt then take and return arrays)"""
        self.mul, self.intt, self.array_mode = mul, intt, True
        return self

    # coefficient vectors are lists of ints (CPU tests, oracle products) or (n, 2) uint64 arrays (the GPU path: no per-element
    # packing on the way to the library); only structural operations happen here, plus one subtraction on the L results
    @staticmethod
    def _zeros(like, k):
        return np.zeros((k, 2), dtype=np.uint64) if isinstance(like, np.ndarray) else [0] * k

    @classmethod
    def _fit(cls, a, k):
        a = a[:k]
        if len(a) == k:
            return a
        pad = cls._zeros(a, k - len(a))
        return np.concatenate([a.reshape(-1, 2), pad]) if isinstance(a, np.ndarray) else list(a) + pad

    @staticmethod
    def _rev(a):
        return np.ascontiguousarray(a[::-1]) if isinstance(a, np.ndarray) else a[::-1]

    def _mul_trunc(self, a, b, k):
        return self._fit(self.mul(a, b), k)

    def __call__(self, values):
        L, N, m = self.L, self.N, self.m
        assert len(values) == L
        if L == 1:
            return [int(values[0]) % P]
        as_array = self.array_mode
        v = pack([int(x) for x in values]) if as_array else list(values)
        pt = self.intt(self._fit(v, N))                        # degree < N, right on the prefix (and zero on the rest)
        if m == 0:
            return unpack(pt) if as_array else list(pt)
        Z = self.Z_arr if as_array else self.Z
        g = self.inv_arr if as_array else self.inv_rev_Z
        q_rev = self._mul_trunc(self._fit(self._rev(pt), m), g, m)   # rev(quotient) = rev(pt) * rev(Z)^-1 mod x^m
        zq = self._mul_trunc(Z, self._rev(q_rev), N)
        a, b = (unpack(pt[:L]), unpack(zq[:L])) if as_array else (pt[:L], zq[:L])
        return [(x - y) % P for x, y in zip(a, b)]


def zerofier_of(points):
    """prod (X - x) over a handful of points (boundary / transition zerofiers, stark.rs:186-213): exact host arithmetic, the
    polynomial fast_zerofier returns"""
    z = [1]
    for x in points:
        z = [(a - x * b) % P for a, b in zip([0] + z, z + [0])]
    return z


def sample_many(rng, count):
    """count x Field::sample(rng(17)) (stark.rs:293-295, 428-430) as a (count, 2) uint64 array.  os.urandom is drawn in one call -
    any split of OS entropy is as random as another; every other byte source is called once per element, so reproducible streams
    give the same elements as the one-by-one loop of `prove`."""
    import os
    if rng is os.urandom:
        raw = np.frombuffer(rng(17 * count), dtype=np.uint8).reshape(count, 17)[:, 1:]     # sample keeps the last 16 bytes, big-endian
        hi = raw[:, :8].copy().view(">u8").reshape(count).astype(np.uint64)
        lo = raw[:, 8:].copy().view(">u8").reshape(count).astype(np.uint64)
        p_hi, p_lo = np.uint64(P >> 64), np.uint64(P & ((1 << 64) - 1))
        ge = (hi > p_hi) | ((hi == p_hi) & (lo >= p_lo))                                   # value >= p: subtract p once (p > 2^127)
        borrow = (lo < p_lo) & ge
        lo = np.where(ge, lo - p_lo, lo)
        hi = np.where(ge, hi - p_hi - borrow.astype(np.uint64), hi)
        return np.stack([lo, hi], axis=1)
    return pack([Stark.sample(rng(17)) for _ in range(count)])


def lagrange_interpolate(domain, values):
    """Polynomial::interpolate_domain (polynomial.rs:123-148) for a handful of points (the boundary conditions):
    exact host arithmetic; the interpolant is unique, so it is the polynomial fast_interpolate_domain returns."""
    acc = [0] * len(domain)
    for i, xi in enumerate(domain):
        num, den = [values[i] % P], 1
        for j, xj in enumerate(domain):
            if i != j:
                num = [(a - xj * b) % P for a, b in zip([0] + num, num + [0])]        # num * (x - xj)
                den = den * (xi - xj) % P
        inv = pow(den, P - 2, P)
        for k, c in enumerate(num):
            acc[k] = (acc[k] + c * inv) % P
    return acc


def deterministic_rng(seed: bytes):
    """A reproducible stand-in for thread_rng().fill_bytes (stark.rs:283, 428): call k returns SHAKE256(seed || u64_be(k)).
    Production callers pass os.urandom."""
    import hashlib
    state = {"n": 0}

    def fill(n):
        out = hashlib.shake_256(seed + state["n"].to_bytes(8, "big")).digest(n)
        state["n"] += 1
        return out
    return fill


def _bit_count(v):
    """BitIter::from(v).count() (utils/bit_iter.rs): bits from the highest set bit down; 0 counts one."""
    return max(v.bit_length(), 1)


def _degree(p):
    d = None
    for i, c in enumerate(p):
        if c:
            d = i
    return d


class Stark:
    def __init__(self, expansion_factor, num_collinearity_checks, security_level, num_registers, num_cycles,
                 transition_constraints_degree, ctx=None):
        """Stark::new stark.rs:71-113"""
        assert _bit_count(P) >= security_level
        assert expansion_factor & (expansion_factor - 1) == 0, "expansion_factor must be a power of 2"
        assert expansion_factor >= 4, "expansion_factor must be at least 4"
        assert num_collinearity_checks * 2 >= security_level
        self.ctx = ctx or default_context()
        self.field = Field()
        self.expansion_factor = expansion_factor
        self.num_registers = num_registers
        self.original_trace_length = num_cycles
        self.num_randomizers = 4 * num_collinearity_checks
        randomized_trace_length = num_cycles + self.num_randomizers
        self.omicron_domain_length = 1 << _bit_count(randomized_trace_length * transition_constraints_degree)
        self.fri_domain_length = self.omicron_domain_length * expansion_factor
        self.generator = self.field.generator()
        self.omega = self.field.primitive_nth_root(self.fri_domain_length)
        self.omicron = self.field.primitive_nth_root(self.omicron_domain_length)
        self.fri = FRI(self.generator, self.omega, self.fri_domain_length, expansion_factor, num_collinearity_checks, ctx=self.ctx)
        self._cache = {}          # what depends on the AIR's shape only: zerofiers, the trace-domain interpolator

    # ---- degree bookkeeping: walks dictionary KEYS, zero coefficients included (stark.rs:115-184) ----
    def transition_degree_bounds(self, transition_constraints):
        points_degree = [1] + [self.original_trace_length + self.num_randomizers - 1] * (2 * self.num_registers)
        res = []
        for a in transition_constraints:
            d = getattr(a, "dictionary", a)
            assert d, "cannot calculate max on empty vec a"
            res.append(max(sum(r * l for r, l in zip(points_degree, k)) for k in d))
        return res

    def transition_quotient_degree_bounds(self, transition_constraints):
        return [d - (self.original_trace_length - 1) for d in self.transition_degree_bounds(transition_constraints)]

    def max_degree(self, transition_constraints):
        return (1 << _bit_count(max(self.transition_degree_bounds(transition_constraints)))) - 1

    def _omicron_pow(self, e):
        return pow(self.omicron, e, P)

    def _prefix_zerofier(self, length):
        """zerofier of omicron^0 .. omicron^(length-1) (fast_zerofier, ntt_arithmetics.rs:66-108); depends on the shape only"""
        key = ("zerofier", length)
        if key not in self._cache:
            self._cache[key] = fft.fast_zerofier(self.omicron, self.omicron_domain_length, [self._omicron_pow(i) for i in range(length)], self.ctx)
        return self._cache[key]

    def transition_zerofier(self):                                        # stark.rs:186-194
        return self._prefix_zerofier(self.original_trace_length - 1)

    def boundary_zerofiers(self, boundary):                               # stark.rs:196-213
        out = []
        for s in range(self.num_registers):
            key = ("boundary_zerofier",) + tuple(c for c, r, _ in boundary if r == s)
            if key not in self._cache:
                self._cache[key] = fft.fast_zerofier(self.omicron, self.omicron_domain_length, [self._omicron_pow(c) for c in key[1:]], self.ctx)
            out.append(self._cache[key])
        return out

    def boundary_interpolants(self, boundary):                            # stark.rs:215-243
        out = []
        for s in range(self.num_registers):
            domain = [self._omicron_pow(c) for c, r, _ in boundary if r == s]
            values = [v for _, r, v in boundary if r == s]
            if len(domain) <= 8:
                out.append(lagrange_interpolate(domain, values))
            else:
                out.append(fft.fast_interpolate_domain(self.omicron, self.omicron_domain_length, domain, values, self.ctx))
        return out

    def trace_interpolator(self, length):
        """stark.rs:305-326 for every column of a trace of `length` rows (see PrefixInterpolator)"""
        key = ("interpolator", length)
        if key not in self._cache:
            ctx, n = self.ctx, self.fri_domain_length
            mul = lambda a, b: fft.fast_multiply(self.omega, n, a, b, ctx)           # noqa: E731  products of degree < 2*omicron_domain_length <= n
            intt = lambda v: fft.intt(self.omicron, v, ctx)                           # noqa: E731
            self._cache[key] = PrefixInterpolator(length, self.omicron_domain_length, self._prefix_zerofier(length), mul, intt).use_arrays(mul, intt)
        return self._cache[key]

    @staticmethod
    def sample(data):
        """Field::sample field.rs:87-99: the wrapping shift-xor fold keeps the big-endian value of the LAST 16 bytes; then % p"""
        return int.from_bytes(bytes(data)[-16:], "big") % P

    def sample_weights(self, number, randomness):                         # stark.rs:260-274 (all weights equal: SURVEY.md A.6)
        return [self.sample(bytes(i) + randomness) for i in range(number)]

    def prove(self, trace, transition_constraints, boundary, proof_stream, rng, check_degrees=True, lockstep=True):
        """Returns the proof bytes (proof_stream.digest()).  proof_stream: the library's IndependentProofStream /
        SignatureProofStream.  lockstep=True (default) runs the batched pipeline with a batch of one - every stage a single
        device call (0.9 ms + the byte source per RPSSS signature); lockstep=False keeps the reference's call structure
        (fast_coset_divide per register, one LDE / commit per codeword: 4.8 ms).  Same draws from `rng`, same bytes."""
        if lockstep:
            return self.prove_batch([trace], transition_constraints, [boundary], [proof_stream], [rng], check_degrees=check_degrees)[0]
        import torch
        ctx, n, nr = self.ctx, self.fri_domain_length, self.num_registers
        dev = torch.device("cuda", ctx.device)
        trace = [list(row) for row in trace]
        for _ in range(self.num_randomizers):                                         # stark.rs:286-301
            trace.append([self.sample(rng(17)) for _ in range(nr)])
        interpolate = self.trace_interpolator(len(trace))                            # stark.rs:303-326
        trace_polynomials = [interpolate([row[s] for row in trace]) for s in range(nr)]
        interpolants = self.boundary_interpolants(boundary)
        zerofiers = self.boundary_zerofiers(boundary)
        boundary_quotients = []
        for s in range(nr):                                                           # stark.rs:331-360
            num = _psub(trace_polynomials[s], interpolants[s])
            boundary_quotients.append(fft.fast_coset_divide(self.omicron, self.omicron_domain_length, self.generator, num, zerofiers[s], ctx))
        # committed codewords live in ONE device buffer: the registers' boundary quotients, then the randomizer
        cws = torch.empty((nr + 1, n, 2), dtype=torch.int64, device=dev)
        trees = []
        try:
            for s in range(nr):                                                       # stark.rs:366-386
                self._lde_into(boundary_quotients[s], cws[s])
                trees.append(MerkleTree(cws[s], ctx))
                proof_stream.push((ROOT, trees[-1].root()))
            tz = self.transition_zerofier()
            tcd = self._air_shape(transition_constraints)[0]
            randomizer_polynomial = [self.sample(rng(17)) for _ in range(tcd + 1)]   # stark.rs:424-432
            self._lde_into(randomizer_polynomial, cws[nr])
            trees.append(MerkleTree(cws[nr], ctx))
            proof_stream.push((ROOT, trees[-1].root()))                               # stark.rs:441-445
            nc = len(transition_constraints)
            weights = self.sample_weights(1 + 2 * nc + 2 * nr, proof_stream.fiat_shamir_prover(PROOF_BYTES))
            tcd, tq_bounds, flat = self._air_shape(transition_constraints)
            bq_bounds = [len(trace) - 1 - _degree(bz) for bz in zerofiers]            # stark.rs:245-258
            shifts = [tcd - b for b in tq_bounds] + [tcd - b for b in bq_bounds]
            # stark.rs:388-519 in evaluation form: quotients, x^shift products, weighted sum - one kernel, one codeword out
            res = air_combination(self.generator, self.omega, n, self.expansion_factor, flat, zerofiers, interpolants,
                                  tz, weights, shifts, cws[:nr], cws[nr], want_quotients=check_degrees, ctx=ctx)
            if check_degrees:                                                         # stark.rs:451-464
                combined, tq_cws = res
                for j in range(nc):
                    deg = self._coset_degree(tq_cws[j])
                    if deg is None:
                        raise ValueError("Failed to get degree of transition quotient")
                    if deg != tq_bounds[j]:
                        raise ValueError("transition quotient degrees do not match with expectation")
            else:
                combined = res
            indices = self.fri.prove(combined, proof_stream)                          # stark.rs:522
            dup = list(indices) + [(i + self.expansion_factor) % n for i in indices]  # stark.rs:524-543
            quad = sorted(dup + [(i + n // 2) % n for i in dup])
            for t in trees:                                                           # stark.rs:546-560
                t.open_into(quad, proof_stream)
            return proof_stream.digest()
        finally:
            try:                                   # torch's allocator does not know the context's own stream: drain it before the tensors go
                ctx.sync()
            except Exception:                      # noqa: BLE001 - the original error is the one to report
                pass
            for t in trees:
                t.close()

    # ---- batches of instances of one AIR in lockstep (csrc/stark.cu, csrc/air.cu, csrc/batch.cu) --------------------------
    def prove_batch(self, traces, transition_constraints, boundaries, proof_streams, rngs, check_degrees=True, return_bytes=True, native=True):
        """Stark::prove for B instances of the same AIR (same constraint list, same boundary POSITIONS; values, traces, documents and
        randomness differ), every launch carrying all of them: trace interpolation + LDE (zkb_trace_lde_batch), boundary quotients in
        evaluation form (zkb_air_boundary_quotients), commits (zkb_merkle_build_batch), the evaluation-form middle (zkb_air_combine),
        FRI (zkb_fri_prove_batch) and the openings (zkb_merkle_open_ps_batch).  Per instance the proof bytes are those of `prove`.
        traces: per instance a list of rows of ints, or a packed (rows, registers, 2) uint64 array.  rngs: one byte source per
        instance (os.urandom draws in bulk).  Returns the list of proofs; with return_bytes=False the proofs stay in the proof streams
        (host memory, read them with .digest()) and their lengths are returned."""
        import ctypes
        import torch
        from . import _lib
        from .context import le16
        ctx, lib, n, nr, B = self.ctx, self.ctx.lib, self.fri_domain_length, self.num_registers, len(traces)
        assert B == len(boundaries) == len(proof_streams) == len(rngs) and B >= 1
        import os
        if native and not os.environ.get("ZKB_STAGED_PROVER"):
            # the whole call sequence below as ONE C call (csrc/prover.cu, zkb_stark_prove_batch): same bytes, no Python between the stages
            return self._prove_batch_native(traces, transition_constraints, boundaries, proof_streams, rngs, check_degrees, return_bytes)
        dev = torch.device("cuda", ctx.device)
        L = len(traces[0]) + self.num_randomizers
        positions = tuple((c, r) for c, r, _ in boundaries[0])
        assert all(tuple((c, r) for c, r, _ in b) == positions for b in boundaries), "a batch shares the boundary positions"
        tcd, tq_bounds, _ = self._air_shape(transition_constraints)
        nc = len(transition_constraints)
        air, bq_bounds = self._air_handle(transition_constraints, positions, L)
        # randomized traces, register-major so that column s*B + b of the batched calls is plane s, instance b
        values = np.zeros((nr, B, L, 2), dtype=np.uint64)
        rnd_polys = np.zeros((B, tcd + 1, 2), dtype=np.uint64)
        t0 = len(traces[0])
        import os
        n_tr, n_rp = self.num_randomizers * nr, tcd + 1
        bulk = sample_many(os.urandom, B * (n_tr + n_rp)).reshape(B, n_tr + n_rp, 2) if all(r is os.urandom for r in rngs) else None
        for b in range(B):
            assert len(traces[b]) == t0
            tr = traces[b]
            if isinstance(tr, np.ndarray):
                values[:, b, :t0] = tr.reshape(t0, nr, 2).transpose(1, 0, 2)
            else:
                values[:, b, :t0] = pack([row[s] for s in range(nr) for row in tr]).reshape(nr, t0, 2)
            draws = bulk[b, :n_tr] if bulk is not None else sample_many(rngs[b], n_tr)   # stark.rs:286-301: row by row, register by register
            values[:, b, t0:] = draws.reshape(self.num_randomizers, nr, 2).transpose(1, 0, 2)
        tcw = torch.empty((nr, B, n, 2), dtype=torch.int64, device=dev)              # trace codewords
        cws = torch.empty((nr + 2, B, n, 2), dtype=torch.int64, device=dev)          # planes: boundary quotients | randomizer | combination
        ctx.check(lib.zkb_trace_lde_batch(ctx.h, le16(self.omicron), self.omicron_domain_length, L, le16(self.omega), n, le16(self.generator),
                                          values.ctypes.data, L, nr * B, tcw.data_ptr(), n, None))          # stark.rs:303-326 + LDE
        ilen = max(len([1 for _, r in positions if r == s]) for s in range(nr)) or 1
        interp = np.zeros((B, nr, ilen, 2), dtype=np.uint64)
        basis = self._lagrange_basis(positions)
        for b in range(B):                                                            # stark.rs:215-243 (a handful of points each)
            for s in range(nr):
                vals = [v for _, r, v in boundaries[b] if r == s]
                if vals:                                                              # sum_i value_i * basis_i: the unique interpolant
                    poly = [sum(v * bp[k] for v, bp in zip(vals, basis[s])) % P for k in range(len(vals))]
                    interp[b, s, :len(poly)] = pack(poly)
        ctx.check(lib.zkb_air_set_interpolants(air, B, interp.ctypes.data, ilen))
        ctx.check(lib.zkb_air_boundary_quotients(air, B, tcw.data_ptr(), B * n, n, cws.data_ptr(), B * n, n))   # stark.rs:331-360
        for b in range(B):
            rnd_polys[b] = bulk[b, n_tr:] if bulk is not None else sample_many(rngs[b], n_rp)   # stark.rs:424-432
        ctx.check(lib.zkb_coset_lde_batch(ctx.h, le16(self.omega), n, le16(self.generator), rnd_polys.ctypes.data, tcd + 1, tcd + 1,
                                          cws[nr].data_ptr(), n, B))
        K = nr + 1
        trees = (ctypes.c_void_p * (K * B))()
        handles = [p.h.value for p in proof_streams]
        ps_arr = (ctypes.c_void_p * B)(*handles)
        ps_of_tree = (ctypes.c_void_p * (K * B))(*(handles * K))                      # tree t*B + b belongs to proof b
        ctx.check(lib.zkb_merkle_build_batch(ctx.h, cws.data_ptr(), n, n, K * B, trees, ps_of_tree))   # stark.rs:366-386, 441-445
        try:
            nw = 1 + 2 * nc + 2 * nr
            weights = np.zeros((B, nw, 2), dtype=np.uint64)
            for b, p in enumerate(proof_streams):                                     # stark.rs:447-450 (all weights of a proof are equal, A.6)
                weights[b, :] = pack([self.sample(p.fiat_shamir_prover(PROOF_BYTES))])[0]
            tq = torch.empty((B, nc, n, 2), dtype=torch.int64, device=dev) if check_degrees else None
            ctx.check(lib.zkb_air_combine(air, B, weights.ctypes.data_as(_lib.c_u8p), cws.data_ptr(), B * n, n, cws[nr].data_ptr(), n,
                                          cws[nr + 1].data_ptr(), n, tq.data_ptr() if check_degrees else None, nc * n))   # stark.rs:388-519
            if check_degrees:                                                         # stark.rs:451-464
                degs = (ctypes.c_int64 * (B * nc))()
                ctx.check(lib.zkb_coset_degree_batch(ctx.h, le16(self.omega), tq.data_ptr(), n, n, B * nc, degs))
                for b in range(B):
                    got = list(degs[b * nc:(b + 1) * nc])
                    if any(d < 0 for d in got):
                        raise ValueError("Failed to get degree of transition quotient")
                    if got != tq_bounds:
                        raise ValueError("transition quotient degrees do not match with expectation")
            top = np.empty((B, self.fri.num_colinearity_tests), dtype=np.uint64)
            ctx.check(lib.zkb_fri_prove_batch(ctx.h, ctypes.byref(self.fri.params), cws[nr + 1].data_ptr(), n, n, B, ps_arr,
                                              top.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))               # stark.rs:522
            nn, ef = np.uint64(n), np.uint64(self.expansion_factor)
            dup = np.concatenate([top, (top + ef) % nn], axis=1)                      # stark.rs:524-543
            quad = np.sort(np.concatenate([dup, (dup + nn // np.uint64(2)) % nn], axis=1), axis=1)
            k = quad.shape[1]
            idx = np.ascontiguousarray(np.broadcast_to(quad[None, :, :], (K, B, k)))
            ctx.check(lib.zkb_merkle_open_ps_batch(trees, K * B, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k, ps_of_tree))   # stark.rs:546-560
            for p in proof_streams:
                p.objects = None
            if return_bytes:
                return [p.digest() for p in proof_streams]
            return [int(lib.zkb_ps_digest(p.h, None, 0)) for p in proof_streams]
        finally:
            # tcw / cws / tq come from torch's caching allocator, which does not know the context's own stream: nothing queued on that
            # stream may still touch them when they are released (error paths included)
            try:
                ctx.sync()
            except Exception:      # noqa: BLE001 - the original error is the one to report
                pass
            for i in range(K * B - 1, -1, -1):          # tree 0 owns the shared arena: free it last
                if trees[i]:
                    lib.zkb_merkle_free(trees[i])

    def _prove_batch_native(self, traces, transition_constraints, boundaries, proof_streams, rngs, check_degrees, return_bytes):
        import ctypes
        import os
        from . import _lib
        from .context import ZkbError
        ctx, lib, nr, B = self.ctx, self.ctx.lib, self.num_registers, len(traces)
        t0 = len(traces[0])
        positions = tuple((c, r) for c, r, _ in boundaries[0])
        assert all(tuple((c, r) for c, r, _ in b) == positions for b in boundaries), "a batch shares the boundary positions"
        air, _ = self._air_handle(transition_constraints, positions, t0 + self.num_randomizers)
        shape = self._stark_shape(transition_constraints, positions, t0, check_degrees)
        tr = np.empty((B, t0, nr, 2), dtype=np.uint64)
        for b, t in enumerate(traces):
            assert len(t) == t0
            tr[b] = t.reshape(t0, nr, 2) if isinstance(t, np.ndarray) else pack([v for row in t for v in row]).reshape(t0, nr, 2)
        bv = pack([v for bd in boundaries for _, _, v in bd]) if positions else np.zeros((1, 2), dtype=np.uint64)
        n_tr, n_rp = self.num_randomizers * nr, shape.rnd_poly_len
        rnd = None
        if not all(r is os.urandom for r in rngs):       # reproducible byte sources: the same draws, in the same order, as `prove` (stark.rs:286-301, 424-432)
            rnd = np.ascontiguousarray(np.stack([np.concatenate([sample_many(r, n_tr), sample_many(r, n_rp)]) for r in rngs]))
        ps_arr = (ctypes.c_void_p * B)(*[p.h.value for p in proof_streams])
        lens = (ctypes.c_uint64 * B)()
        try:
            ctx.check(lib.zkb_stark_prove_batch(ctx.h, air, ctypes.byref(shape), B, tr.ctypes.data, bv.ctypes.data,
                                                rnd.ctypes.data if rnd is not None else None, ps_arr, lens))
        except ZkbError as e:
            if e.code == -11:                            # ZKB_ERR_DEGREE: the degree check of stark.rs:451-464 (a panic in the reference)
                raise ValueError(str(e)) from e
            raise
        for p in proof_streams:
            p.objects = None
        if return_bytes:
            return [p.digest() for p in proof_streams]
        return [int(x) for x in lens]

    def _stark_shape(self, transition_constraints, positions, trace_length, check_degrees):
        """zkb_stark_shape for (constraints, boundary positions, trace length): everything zkb_stark_prove_batch needs beside the AIR handle"""
        import ctypes
        from . import _lib
        key = ("stark_shape", id(transition_constraints), positions, trace_length, bool(check_degrees))
        if key not in self._cache:
            nr = self.num_registers
            tcd, tq_bounds, (counts, _, _) = self._air_shape(transition_constraints)
            d = _lib.StarkShape()
            d.omicron[:] = list(int(self.omicron).to_bytes(16, "little"))
            d.omicron_order, d.trace_length, d.num_randomizers, d.rnd_poly_len = self.omicron_domain_length, trace_length, self.num_randomizers, tcd + 1
            d.num_registers, d.num_constraints, d.num_boundary = nr, len(counts), len(positions)
            bounds = np.asarray(tq_bounds, dtype=np.int64)
            d.tq_degree_bounds = bounds.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if check_degrees else None
            regs = np.asarray([r for _, r in positions] or [0], dtype=np.uint32)
            d.boundary_register = regs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
            basis = self._lagrange_basis(positions)
            lag = pack([c for s in range(nr) for poly in basis[s] for c in poly] or [0])
            d.lagrange = lag.ctypes.data
            d.fri = self.fri.params
            d.proof_bytes = PROOF_BYTES
            self._cache[key] = (d, bounds, regs, lag, transition_constraints)
        return self._cache[key][0]

    def _lagrange_basis(self, positions):
        """per register the Lagrange basis polynomials of its boundary points (they depend on the positions only)"""
        key = ("lagrange", positions)
        if key not in self._cache:
            out = []
            for s in range(self.num_registers):
                dom = [self._omicron_pow(c) for c, r in positions if r == s]
                out.append([lagrange_interpolate(dom, [1 if j == i else 0 for j in range(len(dom))]) for i in range(len(dom))])
            self._cache[key] = out
        return self._cache[key]

    def _air_handle(self, transition_constraints, positions, trace_length):
        """zkb_air for (constraints, boundary positions): grouped terms + zerofier codewords, created once"""
        import ctypes
        from . import _lib
        from .context import Vec, le16
        key = ("air_handle", id(transition_constraints), positions, trace_length)
        if key not in self._cache:
            nr = self.num_registers
            tcd, tq_bounds, (counts, coefs, exps) = self._air_shape(transition_constraints)
            zerofiers = [zerofier_of([self._omicron_pow(c) for c, r in positions if r == s]) for s in range(nr)]
            tz = zerofier_of([self._omicron_pow(i) for i in range(self.original_trace_length - 1)])
            bq_bounds = [trace_length - 1 - _degree(z) for z in zerofiers]             # stark.rs:245-258
            shifts = np.asarray([tcd - b for b in tq_bounds] + [tcd - b for b in bq_bounds], dtype=np.uint64)
            d = _lib.AirShape()
            d.offset[:] = list(int(self.generator).to_bytes(16, "little"))
            d.omega[:] = list(int(self.omega).to_bytes(16, "little"))
            d.domain_length, d.expansion_factor, d.num_registers, d.num_constraints = self.fri_domain_length, self.expansion_factor, nr, len(counts)
            d.term_counts = counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
            d.coefs = coefs.ctypes.data_as(_lib.c_u8p)
            d.exps = exps.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
            zv = [Vec(z) for z in zerofiers]
            zp = (ctypes.c_void_p * nr)(*[v.ptr for v in zv])
            zl = (ctypes.c_size_t * nr)(*[v.n for v in zv])
            d.boundary_zerofiers, d.boundary_zerofier_lens = zp, zl
            tzv = Vec(tz)
            d.transition_zerofier, d.transition_zerofier_len = tzv.ptr, tzv.n
            d.shifts = shifts.ctypes.data_as(_lib.c_u64p)
            h = ctypes.c_void_p()
            self.ctx.check(self.ctx.lib.zkb_air_create(self.ctx.h, ctypes.byref(d), ctypes.byref(h)))
            self._cache[key] = (h, bq_bounds, transition_constraints)
        return self._cache[key][:2]

    def close(self):
        """release the device tables of the cached AIR handles"""
        for key, val in list(self._cache.items()):
            if key[0] == "air_handle":
                self.ctx.lib.zkb_air_free(val[0])
                del self._cache[key]

    def _air_shape(self, transition_constraints):
        """(max_degree, transition quotient degree bounds, flattened terms) of an AIR: computed once per constraint list
        (the list is kept alive by the cache entry, so its id cannot be recycled)"""
        key = ("air", id(transition_constraints))
        if key not in self._cache:
            from .air import flatten_constraints
            self._cache[key] = (self.max_degree(transition_constraints), self.transition_quotient_degree_bounds(transition_constraints),
                                flatten_constraints(transition_constraints, self.num_registers), transition_constraints)
        return self._cache[key][:3]

    def _lde_into(self, coefficients, out):
        ctx = self.ctx
        from .context import Vec, le16
        v = Vec(list(coefficients))
        ctx.check(ctx.lib.zkb_coset_lde(ctx.h, le16(self.omega), self.fri_domain_length, le16(self.generator),
                                        v.ptr if v.n else None, v.n, out.data_ptr()))

    def _coset_degree(self, codeword):
        """degree of the polynomial whose values on offset*<omega> are `codeword`: iNTT gives c_i * offset^i, and
        offset != 0, so the last non-zero entry is the degree (no un-scaling needed)"""
        coeffs = fft.intt(self.omega, codeword, self.ctx)
        self.ctx.sync()                                   # the context may run on its own stream; torch reads on its current one
        nz = (coeffs != 0).any(dim=1).nonzero()
        return int(nz[-1]) if nz.numel() else None


def _psub(a, b):
    """Polynomial - (polynomial.rs:252-288): a + (-b) with the reference's early returns for zero operands"""
    nb = [(-c) % P for c in b]
    if _degree(a) is None:
        return nb
    if _degree(nb) is None:
        return list(a)
    out = [0] * max(len(a), len(nb))
    for i, c in enumerate(a):
        out[i] = (out[i] + c) % P
    for i, c in enumerate(nb):
        out[i] = (out[i] + c) % P
    return out
