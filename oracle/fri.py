"""FRI prover / verifier (oracle; test infrastructure only).  Restates src/fri.rs.

``commit`` / ``query`` / ``prove`` follow fri.rs:115-248, ``verify`` fri.rs:250-416,
``sample_index(es)`` fri.rs:60-113, ``num_rounds`` fri.rs:40-50.
"""
from . import field as F
from . import ntt as N
from . import merkle as M
from . import proof_stream as PS


def test_colinearity(points) -> bool:
    # src/field/polynomial.rs:122-177: Lagrange-interpolate the 3 points, accept iff
    # the interpolant has degree exactly 1.
    xs = [p[0] for p in points]
    ys = [p[1] for p in points]
    acc = [0] * len(xs)
    for i, xi in enumerate(xs):
        prod = [ys[i]]
        for j, xj in enumerate(xs):
            if i == j:
                continue
            s = F.inv(F.sub(xi, xj))
            nxt = [0] * (len(prod) + 1)
            for k, c in enumerate(prod):                 # prod * (x - xj) * s
                nxt[k] = (nxt[k] - c * xj) % F.P
                nxt[k + 1] = (nxt[k + 1] + c) % F.P
            prod = [c * s % F.P for c in nxt]
        for k, c in enumerate(prod):
            acc[k] = (acc[k] + c) % F.P
    return N.degree(acc) == 1


class FRI:
    def __init__(self, offset, omega, domain_length, expansion_factor, num_colinearity_tests):
        self.offset, self.omega = offset, omega           # fri.rs:23-38
        self.domain_length = domain_length
        self.expansion_factor = expansion_factor
        self.num_colinearity_tests = num_colinearity_tests

    def num_rounds(self):
        n, r = self.domain_length, 0                      # fri.rs:40-50
        while n > self.expansion_factor and n > 4 * self.num_colinearity_tests:
            n //= 2
            r += 1
        return r

    def evaluate_domain(self):
        return [self.offset * F.fpow(self.omega, i) % F.P for i in range(self.domain_length)]

    @staticmethod
    def sample_index(data: bytes, size: int) -> int:
        # fri.rs:60-83: big-endian int of the LAST (bit_index(size)/8 + 1) bytes, % size
        assert size != 0, "modulo zero is impossible"
        nbytes = (size.bit_length() - 1) // 8 + 1
        return int.from_bytes(data[max(0, len(data) - nbytes):], "big") % size

    @classmethod
    def sample_indices(cls, seed: bytes, size, reduced_size, number):
        # fri.rs:85-113: counter = a run of `counter` zero BYTES appended to the seed
        assert number <= 2 * reduced_size, "Not enough entropy in indices with reference to last codeword"
        assert number <= reduced_size, "Cannot sample more indices than available in the last codeword"
        indices, reduced, counter = [], [], 0
        while len(indices) < number:
            idx = cls.sample_index(M.blake2b512(seed + bytes(counter)), size)
            counter += 1
            if idx % reduced_size not in reduced:
                indices.append(idx)
                reduced.append(idx % reduced_size)
        return indices

    @staticmethod
    def fold(codeword, alpha, offset, omega):
        # fri.rs:150-159, formula kept literally
        half = len(codeword) // 2
        two_inv = F.inv(2)
        out = []
        x = offset
        for i in range(half):
            a_by_x = F.div(alpha, x)                       # alpha / (offset * omega^i)
            first = F.mul(F.add(1, a_by_x), codeword[i])
            second = F.mul(F.sub(1, a_by_x), codeword[half + i])
            out.append(F.mul(two_inv, F.add(first, second)))
            x = F.mul(x, omega)
        return out

    def commit(self, codeword, ps):
        # fri.rs:115-172
        omega, offset = self.omega, self.offset
        rounds = self.num_rounds()
        codewords = []
        codeword = list(codeword)
        for r in range(rounds):
            n = len(codeword)
            assert F.fpow(omega, n - 1) == F.inv(omega), \
                "error in commit: omega does not have the right order!"
            ps.push((PS.ROOT, M.commit(codeword)))
            if r == rounds - 1:
                break
            alpha = F.sample(ps.fiat_shamir_prover(PS.PROOF_BYTES))
            codewords.append(codeword)
            codeword = self.fold(codeword, alpha, offset, omega)
            omega = F.mul(omega, omega)
            offset = F.mul(offset, offset)
        ps.push((PS.CODEWORD, list(codeword)))
        codewords.append(codeword)
        return codewords

    def query(self, cur, nxt, indices_c, ps):
        # fri.rs:174-208
        ncc = self.num_colinearity_tests
        ia = list(indices_c)
        ib = [i + len(cur) // 2 for i in indices_c]
        for s in range(ncc):
            ps.push((PS.LEAFS, (cur[ia[s]], cur[ib[s]], nxt[indices_c[s]])))
        for s in range(ncc):
            ps.push((PS.PATH, M.open_(ia[s], cur)))
            ps.push((PS.PATH, M.open_(ib[s], cur)))
            ps.push((PS.PATH, M.open_(indices_c[s], nxt)))
        return ia + ib

    def prove(self, codeword, ps):
        # fri.rs:210-248
        assert self.domain_length == len(codeword), \
            "Length of the domain doesnt match the length of initial codeword"
        codewords = self.commit(codeword, ps)
        top = self.sample_indices(ps.fiat_shamir_prover(PS.PROOF_BYTES),
                                  len(codewords[1]), len(codewords[-1]),
                                  self.num_colinearity_tests)
        indices = list(top)
        for i in range(len(codewords) - 1):
            indices = [j % (len(codewords[i]) // 2) for j in indices]
            self.query(codewords[i], codewords[i + 1], indices, ps)
        return top

    def verify(self, ps, polynomial_values):
        # fri.rs:250-416.  Returns None on success or an error string.
        omega, offset = self.omega, self.offset
        rounds = self.num_rounds()
        ncc = self.num_colinearity_tests
        roots, alphas = [], []
        for _ in range(rounds):
            kind, root = ps.pull()
            assert kind == PS.ROOT
            roots.append(root)
            alphas.append(F.sample(ps.fiat_shamir_verifier(PS.PROOF_BYTES)))
        kind, last = ps.pull()
        assert kind == PS.CODEWORD
        if M.commit(last) != roots[-1]:
            return "last codeword is not well formed"
        deg_bound = len(last) // self.expansion_factor - 1
        last_omega, last_offset = omega, offset
        for _ in range(rounds - 1):
            last_omega = F.mul(last_omega, last_omega)
            last_offset = F.mul(last_offset, last_offset)
        if F.inv(last_omega) != F.fpow(last_omega, len(last) - 1):
            return "omega does not have the right order"
        poly = N.scale(N.intt(last_omega, last), F.inv(last_offset))
        d = N.degree(poly)
        if d is None:
            return "Received none instead of polynomial degree"
        if d > deg_bound:
            return ("last codeword does not correspond to polynomial of low enough degree "
                    "(it is %d but should be <= %d)" % (d, deg_bound))
        if N.ntt(last_omega, N.scale(poly, last_offset)) != list(last):
            return "re-evaluated codeword does not match original"
        top = self.sample_indices(ps.fiat_shamir_verifier(PS.PROOF_BYTES),
                                  self.domain_length >> 1,
                                  self.domain_length >> (rounds - 1), ncc)
        for r in range(rounds - 1):
            half = self.domain_length >> (r + 1)
            ic = [i % half for i in top]
            ia = list(ic)
            ib = [i + half for i in ia]
            aa, bb, cc = [], [], []
            for s in range(ncc):
                kind, (ay, by, cy) = ps.pull()
                assert kind == PS.LEAFS
                aa.append(ay); bb.append(by); cc.append(cy)
                if r == 0:
                    polynomial_values.append((ia[s], ay))
                    polynomial_values.append((ib[s], by))
                ax = F.mul(offset, F.fpow(omega, ia[s]))
                bx = F.mul(offset, F.fpow(omega, ib[s]))
                if not test_colinearity([(ax, ay), (bx, by), (alphas[r], cy)]):
                    return "colinearity check failure"
            for s in range(ncc):
                for which, root, idx, val in (("aa", roots[r], ia[s], aa[s]),
                                              ("bb", roots[r], ib[s], bb[s]),
                                              ("cc", roots[r + 1], ic[s], cc[s])):
                    kind, path = ps.pull()
                    assert kind == PS.PATH
                    if not M.verify(root, idx, path, val):
                        return "Merkle auth path verification failed for " + which
            omega = F.mul(omega, omega)
            offset = F.mul(offset, offset)
        return None
