"""The STARK prover / verifier and the RPSSS signature scheme (oracle; TEST INFRASTRUCTURE ONLY -
see oracle/__init__.py): the CALLERS of the hot path, restated so that BASELINE configs[0]
(Rescue-Prime hash-trace STARK prove + verify at the tutorial parameters) and configs[4]'s
signature proofs can be run bit-exactly with either backend.

Restates src/stark/stark.rs: new :71-113, transition_degree_bounds :115-157,
transition_quotient_degree_bounds :159-167, max_degree :169-184, transition_zerofier :186-194,
boundary_zerofiers :196-213, boundary_interpolants :215-243, boundary_quotient_degree_bounds
:245-258, sample_weights :260-274, prove :276-563, verify :565-770; and src/rpsss.rs:17-87.

`Backend` bundles the hot-path functions Stark calls: the default is the oracle's
own; tests/test_gpu_parity.py swaps in the CUDA-backed mirror (zk_stark_tutor_b200) and asserts
the proof bytes are identical and that this verifier accepts them.  The reference draws its
randomizers from thread_rng (stark.rs:283, 428); here the byte source is a parameter."""
import hashlib

from . import field as F
from . import merkle as M
from . import ntt as N
from . import poly as PL
from . import proof_stream as PS
from .fri import FRI
from .mpoly import MPolynomial, bit_count   # noqa: F401
from .rescue_prime import RescuePrime

P = F.P


class Backend:
    """The reference functions Stark / FRI call on the hot path, oracle versions."""
    name = "oracle"
    fast_zerofier = staticmethod(PL.fast_zerofier)
    fast_interpolate_domain = staticmethod(PL.fast_interpolate_domain)
    fast_coset_divide = staticmethod(N.fast_coset_divide)
    fast_coset_evaluate = staticmethod(N.fast_coset_evaluate)
    fast_multiply = staticmethod(N.fast_multiply)
    commit = staticmethod(M.commit)

    @staticmethod
    def open_many(codeword, indices):
        """[MerkleRoot::open(i, codeword) for i in indices] from ONE tree (the reference rebuilds it per index)."""
        lv = M.tree_levels(codeword)
        out = []
        for index in indices:
            path, i = [], index
            for level in lv[:-1]:
                path.append(level[i ^ 1])
                i >>= 1
            out.append(path)
        return out

    @staticmethod
    def fri(offset, omega, domain_length, expansion_factor, num_colinearity_tests):
        return _OracleFri(offset, omega, domain_length, expansion_factor, num_colinearity_tests)


class _OracleFri(FRI):
    def prove(self, codeword, ps):
        from . import cbind as C, fastfri        # C kernels for the Merkle trees (result-identical, checked in tests)
        top, _, _ = fastfri.prove(self, C.to_arr(codeword), ps)
        return top


def deterministic_rng(seed: bytes):
    """A reproducible stand-in for thread_rng().fill_bytes (stark.rs:283): SHAKE256 stream of `seed`."""
    state = {"n": 0}

    def fill(n):
        out = hashlib.shake_256(seed + state["n"].to_bytes(8, "big")).digest(n)
        state["n"] += 1
        return out
    return fill


class Stark:
    def __init__(self, expansion_factor, num_collinearity_checks, security_level, num_registers, num_cycles,
                 transition_constraints_degree, backend=Backend):
        assert bit_count(P) >= security_level
        assert expansion_factor & (expansion_factor - 1) == 0, "expansion_factor must be a power of 2"
        assert expansion_factor >= 4, "expansion_factor must be at least 4"
        assert num_collinearity_checks * 2 >= security_level
        self.B = backend
        self.expansion_factor = expansion_factor
        self.num_registers = num_registers
        self.original_trace_length = num_cycles
        self.num_randomizers = 4 * num_collinearity_checks
        randomized_trace_length = num_cycles + self.num_randomizers
        self.omicron_domain_length = 1 << bit_count(randomized_trace_length * transition_constraints_degree)
        fri_domain_length = self.omicron_domain_length * expansion_factor
        self.generator = F.GENERATOR
        self.omega = F.primitive_nth_root(fri_domain_length)
        self.omicron = F.primitive_nth_root(self.omicron_domain_length)
        self.omicron_domain = [F.fpow(self.omicron, i) for i in range(self.omicron_domain_length)]
        self.fri = backend.fri(self.generator, self.omega, fri_domain_length, expansion_factor, num_collinearity_checks)
        self.fri_domain_length = fri_domain_length

    # ---- degree bookkeeping (walks dictionary KEYS, zero coefficients included) ----------------
    def transition_degree_bounds(self, transition_constraints):
        points_degree = [1] + [self.original_trace_length + self.num_randomizers - 1] * (2 * self.num_registers)
        res = []
        for a in transition_constraints:
            assert a.dictionary, "cannot calculate max on empty vec a"
            res.append(max(sum(r * l for r, l in zip(points_degree, k)) for k in a.dictionary))
        return res

    def transition_quotient_degree_bounds(self, transition_constraints):
        return [d - (self.original_trace_length - 1) for d in self.transition_degree_bounds(transition_constraints)]

    def max_degree(self, transition_constraints):
        md = max(self.transition_degree_bounds(transition_constraints))
        return (1 << bit_count(md)) - 1

    def transition_zerofier(self):
        return self.B.fast_zerofier(self.omicron, self.omicron_domain_length, self.omicron_domain[0:self.original_trace_length - 1])

    def boundary_zerofiers(self, boundary):
        return [self.B.fast_zerofier(self.omicron, self.omicron_domain_length,
                                     [F.fpow(self.omicron, c) for c, r, _ in boundary if r == s])
                for s in range(self.num_registers)]

    def boundary_interpolants(self, boundary):
        out = []
        for s in range(self.num_registers):
            domain = [F.fpow(self.omicron, c) for c, r, _ in boundary if r == s]
            values = [v for _, r, v in boundary if r == s]
            out.append(self.B.fast_interpolate_domain(self.omicron, self.omicron_domain_length, domain, values))
        return out

    def boundary_quotient_degree_bounds(self, randomized_trace_length, boundary):
        return [randomized_trace_length - 1 - PL.degree(bz) for bz in self.boundary_zerofiers(boundary)]

    @staticmethod
    def sample_weights(number, randomness):
        return [F.sample(bytes(i) + randomness) for i in range(number)]

    # ---- prove ----------------------------------------------------------------------------------
    def prove(self, trace, transition_constraints, boundary, proof_stream, rng):
        B = self.B
        trace = [list(row) for row in trace]
        for _ in range(self.num_randomizers):                                         # stark.rs:286-301
            trace.append([F.sample(rng(17)) for _ in range(self.num_registers)])
        trace_domain = [F.fpow(self.omicron, i) for i in range(len(trace))]
        trace_polynomials = [B.fast_interpolate_domain(self.omicron, self.omicron_domain_length, trace_domain,
                                                       [row[s] for row in trace]) for s in range(self.num_registers)]
        interpolants = self.boundary_interpolants(boundary)
        zerofiers = self.boundary_zerofiers(boundary)
        boundary_quotients = [B.fast_coset_divide(self.omicron, self.omicron_domain_length, self.generator,
                                                  PL.sub(trace_polynomials[s], interpolants[s]), zerofiers[s])
                              for s in range(self.num_registers)]
        boundary_quotient_codewords = []
        for s in range(self.num_registers):                                           # stark.rs:368-386
            cw = B.fast_coset_evaluate(self.omega, self.fri_domain_length, self.generator, boundary_quotients[s])
            proof_stream.push((PS.ROOT, B.commit(cw)))
            boundary_quotient_codewords.append(cw)
        point = [[0, 1]] + trace_polynomials + [N.scale(tp, self.omicron) for tp in trace_polynomials]
        transition_zerofier = self.transition_zerofier()

        def mul(a, b):       # products inside evaluate_symbolic through the backend's NTT multiply when they fit
            da, db = PL.degree(a), PL.degree(b)
            if da is None or db is None:
                return []
            if da + db < self.omicron_domain_length:
                return B.fast_multiply(self.omicron, self.omicron_domain_length, list(a[:da + 1]), list(b[:db + 1]))
            return PL.mul(a, b)
        transition_quotients = []
        for tc in transition_constraints:                                             # stark.rs:401-419
            # evaluate_symbolic: exact arithmetic, so the grouped evaluation gives the reference's
            # polynomial (tests/test_oracle_stark.py checks it against the literal loop)
            tp = tc.evaluate_symbolic_grouped(point, mul=mul)
            transition_quotients.append(B.fast_coset_divide(self.omicron, self.omicron_domain_length, self.generator,
                                                            tp, transition_zerofier))
        tcd = self.max_degree(transition_constraints)
        randomizer_polynomial = [F.sample(rng(17)) for _ in range(tcd + 1)]          # stark.rs:424-432
        randomizer_codeword = B.fast_coset_evaluate(self.omega, self.fri_domain_length, self.generator, randomizer_polynomial)
        proof_stream.push((PS.ROOT, B.commit(randomizer_codeword)))
        weights = self.sample_weights(1 + 2 * len(transition_quotients) + 2 * len(boundary_quotients),
                                      proof_stream.fiat_shamir_prover(PS.PROOF_BYTES))
        tq_bounds = self.transition_quotient_degree_bounds(transition_constraints)
        degs = [PL.degree(tq) for tq in transition_quotients]
        if any(d is None for d in degs):
            raise ValueError("Failed to get degree of transition quotient")
        if degs != tq_bounds:
            raise ValueError("transition quotient degrees do not match with expectation")
        x_pow = lambda k: [0] * k + [1]                                               # noqa: E731
        terms = [randomizer_polynomial]
        for i, tq in enumerate(transition_quotients):
            terms.append(tq)
            terms.append(B.fast_multiply(self.omicron, self.omicron_domain_length, x_pow(tcd - tq_bounds[i]), tq))
        bq_bounds = self.boundary_quotient_degree_bounds(len(trace), boundary)
        for i, bq in enumerate(boundary_quotients):
            terms.append(bq)
            terms.append(B.fast_multiply(self.omicron, self.omicron_domain_length, x_pow(tcd - bq_bounds[i]), bq))
        combination = None
        for w, term in zip(weights, terms):                                           # stark.rs:502-512
            t = PL.mul([w], term)
            combination = t if combination is None else PL.add(combination, t)
        combined_codeword = B.fast_coset_evaluate(self.omega, self.fri_domain_length, self.generator, combination)
        indices = self.fri.prove(combined_codeword, proof_stream)
        n = self.fri_domain_length
        dup = list(indices) + [(i + self.expansion_factor) % n for i in indices]
        quad = sorted(dup + [(i + n // 2) % n for i in dup])
        for cw in boundary_quotient_codewords + [randomizer_codeword]:                # stark.rs:546-560
            for i, path in zip(quad, B.open_many(cw, quad)):
                proof_stream.push((PS.VALUE, cw[i]))
                proof_stream.push((PS.PATH, path))
        return proof_stream.digest()

    # ---- verify (host code in the reference as well; always the oracle's arithmetic) ------------
    def verify(self, transition_constraints, boundary, proof_stream):
        """Returns None on success or the reference's error string."""
        original_trace_length = 1 + max(c for c, _, _ in boundary)
        randomized_trace_length = original_trace_length + self.num_randomizers
        bq_roots = []
        for _ in range(self.num_registers):
            kind, root = proof_stream.pull()
            assert kind == PS.ROOT
            bq_roots.append(root)
        kind, randomizer_root = proof_stream.pull()
        assert kind == PS.ROOT
        interpolants = [PL.fast_interpolate_domain(self.omicron, self.omicron_domain_length,
                                                   [F.fpow(self.omicron, c) for c, r, _ in boundary if r == s],
                                                   [v for _, r, v in boundary if r == s]) for s in range(self.num_registers)]
        weights = self.sample_weights(1 + 2 * len(transition_constraints) + 2 * len(interpolants),
                                      proof_stream.fiat_shamir_verifier(PS.PROOF_BYTES))
        points = []
        ofri = FRI(self.generator, self.omega, self.fri_domain_length, self.expansion_factor, self.fri.num_colinearity_tests)
        err = ofri.verify(proof_stream, points)
        if err is not None:
            return "FRI verification failed: " + err
        points.sort(key=lambda p: p[0])
        indices = [p[0] for p in points]
        values = [p[1] for p in points]
        n = self.fri_domain_length
        duplicated = sorted(indices + [(i + self.expansion_factor) % n for i in indices])
        leafs = []
        for bqr in bq_roots:
            d = {}
            for i in duplicated:
                kind, leaf = proof_stream.pull()
                assert kind == PS.VALUE
                kind, path = proof_stream.pull()
                assert kind == PS.PATH
                if not M.verify(bqr, i, path, leaf):
                    return "Boundary quotient root %d is not verified" % i
                d[i] = leaf
            leafs.append(d)
        randomizers = {}
        for i in duplicated:
            kind, leaf = proof_stream.pull()
            assert kind == PS.VALUE
            kind, path = proof_stream.pull()
            assert kind == PS.PATH
            if not M.verify(randomizer_root, i, path, leaf):
                return "Randomizer leaf %d not verified" % i
            randomizers[i] = leaf
        zerofiers = [PL.fast_zerofier(self.omicron, self.omicron_domain_length,
                                      [F.fpow(self.omicron, c) for c, r, _ in boundary if r == s]) for s in range(self.num_registers)]
        tz = PL.fast_zerofier(self.omicron, self.omicron_domain_length, self.omicron_domain[0:self.original_trace_length - 1])
        tcd = self.max_degree(transition_constraints)
        tq_bounds = self.transition_quotient_degree_bounds(transition_constraints)
        bq_bounds = [randomized_trace_length - 1 - PL.degree(bz) for bz in zerofiers]
        for index_i, cur in enumerate(indices):
            x_cur = F.mul(self.generator, F.fpow(self.omega, cur))
            nxt = (cur + self.expansion_factor) % n
            x_nxt = F.mul(self.generator, F.fpow(self.omega, nxt))
            t_cur, t_nxt = [], []
            for s in range(self.num_registers):
                t_cur.append((leafs[s][cur] * PL.evaluate(zerofiers[s], x_cur) + PL.evaluate(interpolants[s], x_cur)) % P)
                t_nxt.append((leafs[s][nxt] * PL.evaluate(zerofiers[s], x_nxt) + PL.evaluate(interpolants[s], x_nxt)) % P)
            point = [x_cur] + t_cur + t_nxt
            tcv = [tc.evaluate(point) for tc in transition_constraints]
            terms = [randomizers[cur]]
            tzv = PL.evaluate(tz, x_cur)
            for s, v in enumerate(tcv):
                q = F.div(v, tzv)
                terms.append(q)
                terms.append(q * F.fpow(x_cur, tcd - tq_bounds[s]) % P)
            for s in range(self.num_registers):
                bqv = leafs[s][cur]
                terms.append(bqv)
                terms.append(bqv * F.fpow(x_cur, tcd - bq_bounds[s]) % P)
            combination = sum(t * w for t, w in zip(terms, weights)) % P
            if combination != values[index_i]:
                return "Combination doesn't match with polynomial value"
        return None


class RPSSS:
    """src/rpsss.rs:17-87 - the Rescue-Prime STARK signature scheme."""

    def __init__(self, expansion_factor=4, num_collinearity_checks=64, security_level=128, transition_constraints_degree=3,
                 backend=Backend):
        self.rp = RescuePrime(2, 1, security_level, 27, interpolate=backend.fast_interpolate_domain)
        self.stark = Stark(expansion_factor, num_collinearity_checks, security_level, self.rp.m, self.rp.N + 1,
                           transition_constraints_degree, backend)
        self._tc = None

    def transition_constraints(self):
        if self._tc is None:
            self._tc = self.rp.transition_constraints(self.stark.omicron, self.stark.omicron_domain_length)
        return self._tc

    def keygen(self, rng):
        sk = F.sample(rng(17))
        return sk, self.rp.hash(sk)

    def sign(self, sk, document, rng, make_stream=PS.SignatureProofStream):
        out = self.rp.hash(sk)
        return self.stark.prove(self.rp.trace(sk), self.transition_constraints(), self.rp.boundary_constraints(out),
                                make_stream(document), rng)

    def verify(self, pk, document, signature):
        sps = PS.SignatureProofStream(document, PS.parse(signature))
        return self.stark.verify(self.transition_constraints(), self.rp.boundary_constraints(pk), sps)
