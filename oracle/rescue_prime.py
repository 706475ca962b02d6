"""Rescue-Prime (oracle; TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Restates src/rescue_prime/rescue_prime.rs: the permutation :51-103, new :106-129, get_mds
:131-151, get_round_constants :153-185, hash :188-195, trace :197-207,
round_constants_polynomials :209-247, transition_constraints :249-287, boundary_constraints
:289-294.  Pinned by the reference's own known answers (rescue_prime.rs:298-331: alpha,
alpha_inv, MDS, MDS_inv, all 108 round constants, hash(1), hash(5732...)) in
tests/test_oracle_stark.py."""
import math

from . import field as F
from . import poly as PL
from . import proof_stream as PS
from .mpoly import MPolynomial, bit_count, inverse, rref, transpose

P = F.P


def smallest_generator():
    # field.rs:46-56
    k = 3
    while math.gcd(k, P - 1) != 1:
        k += 1
    return k


class RescuePrime:
    def __init__(self, m=2, capacity=1, security_level=128, N=27, interpolate=None):
        g = smallest_generator()
        self.m, self.capacity, self.N = m, capacity, N
        self.alpha = g
        self.alpha_inv = F.inv(F.neg(g))                          # rescue_prime.rs:124 (= 1/alpha mod p-1 for this p)
        self.MDS = self.get_mds(g, m)
        self.MDS_inv = inverse(self.MDS)
        self.round_constants = self.get_round_constants(m, capacity, security_level, N)
        self.interpolate = interpolate or PL.fast_interpolate_domain

    @staticmethod
    def get_mds(g, m):
        matrix = [[F.fpow(g, i * j) for j in range(2 * m)] for i in range(m)]
        rref(matrix)
        return transpose([row[m:] for row in matrix])

    @staticmethod
    def get_round_constants(m, capacity, security_level, N):
        bytes_per_int = (bit_count(P) + 7) // 8 + 1
        num_bytes = bytes_per_int * 2 * m * N
        seed = ("Rescue-XLIX(%d,%d,%d,%d)" % (P, m, capacity, security_level)).encode()
        data = PS.shake256(seed, num_bytes)
        out = []
        for i in range(2 * m * N):
            chunk = data[bytes_per_int * i: bytes_per_int * (i + 1)]
            acc = 0
            for j, b in enumerate(chunk):
                acc = (acc + F.fpow(256, j) * b) % P
            out.append(acc)
        return out

    def _round(self, state, r):
        m, rc = self.m, self.round_constants
        s = [F.fpow(x, self.alpha) for x in state]
        acc = [0] * m
        for i, x in enumerate(s):
            for j in range(m):
                acc[j] = (acc[j] + self.MDS[j][i] * x) % P
        s = [(x + rc[2 * r * m + i]) % P for i, x in enumerate(acc)]
        s = [F.fpow(x, self.alpha_inv) for x in s]
        acc = [0] * m
        for i, x in enumerate(s):
            for j in range(m):
                acc[j] = (acc[j] + self.MDS[j][i] * x) % P
        return [(x + rc[2 * r * m + m + i]) % P for i, x in enumerate(acc)]

    def trace(self, input_element):
        state = [input_element] + [0] * (self.m - self.capacity)
        out = [list(state)]
        for r in range(self.N):
            state = self._round(state, r)
            out.append(list(state))
        return out

    def hash(self, input_element):
        return self.trace(input_element)[-1][0]

    def round_constants_polynomials(self, omicron, omicron_domain_length):
        domain = [F.fpow(omicron, r) for r in range(self.N)]
        left, right = [], []
        for i in range(self.m):
            values = [self.round_constants[2 * r * self.m + i] for r in range(self.N)]
            left.append(MPolynomial.lift(self.interpolate(omicron, omicron_domain_length, domain, values), 0))
        for i in range(self.m):
            values = [self.round_constants[2 * r * self.m + self.m + i] for r in range(self.N)]
            right.append(MPolynomial.lift(self.interpolate(omicron, omicron_domain_length, domain, values), 0))
        return left, right

    def transition_constraints(self, omicron, omicron_domain_length):
        first_step, second_step = self.round_constants_polynomials(omicron, omicron_domain_length)
        variables = MPolynomial.variables(1 + 2 * self.m)
        previous_state = variables[1:1 + self.m]
        next_state = variables[1 + self.m:1 + 2 * self.m]
        air = []
        for i in range(self.m):
            lhs = None
            for k in range(self.m):
                t = MPolynomial.constant(self.MDS[i][k]) * (previous_state[k] ** self.alpha)
                lhs = t if lhs is None else lhs + t
            lhs = lhs + first_step[i]
            rhs = None
            for k in range(self.m):
                t = MPolynomial.constant(self.MDS_inv[i][k]) * (next_state[k] - second_step[k])
                rhs = t if rhs is None else rhs + t
            rhs = rhs ** self.alpha
            air.append(lhs - rhs)
        return air

    def boundary_constraints(self, output_element):
        return [(0, 1, 0), (self.N, 0, output_element)]
