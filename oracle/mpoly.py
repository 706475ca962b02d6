"""Multivariate polynomials and the small matrix helpers of the reference (oracle; TEST
INFRASTRUCTURE ONLY - see oracle/__init__.py).

Restates src/m_polynomial.rs (dictionary {exponent vector -> coefficient}; constant :38-45,
variables :50-64, lift :66-84, is_zero :86-95, evaluate :97-126, evaluate_symbolic :128-142,
neg :171-183, add :185-232, sub :234-239, mul :241-281, pow :284-315) and
src/utils/matrix.rs (rref :5-49, transpose :52-65, inverse :67-110).

The dictionary semantics are kept literally - keys of different lengths, no pruning of
zero coefficients - because Stark::transition_degree_bounds (stark.rs:115-157) walks the
KEYS, zero-coefficient entries included."""
from . import field as F
from . import poly as PL

P = F.P


def _bits_msb_first(v):
    # utils/bit_iter.rs: From<u128> + Iterator: bits from the highest set bit down to bit 0; 0 -> one `false`
    if v == 0:
        return [False]
    return [bool((v >> i) & 1) for i in range(v.bit_length() - 1, -1, -1)]


def bit_count(v):
    """BitIter::from(v).count()"""
    return len(_bits_msb_first(v))


class MPolynomial:
    def __init__(self, dictionary=None):
        self.dictionary = dict(dictionary or {})          # {tuple of exponents: coefficient}

    @staticmethod
    def zero():
        return MPolynomial({})

    @staticmethod
    def constant(element):
        return MPolynomial({(0,): element % P})

    @staticmethod
    def variables(num_variables):
        out = []
        for i in range(num_variables):
            e = [0] * i + [1]
            e += [0] * (num_variables - len(e))
            out.append(MPolynomial({tuple(e): 1}))
        return out

    @staticmethod
    def lift(polynomial, variable_index):
        acc = MPolynomial.zero()
        if PL.degree(polynomial) is None:
            return acc
        x = MPolynomial.variables(variable_index + 1)[-1]
        for i, el in enumerate(polynomial):
            acc = acc + MPolynomial.constant(el) * (x ** i)
        return acc

    def is_zero(self):
        return not any(v % P != 0 for v in self.dictionary.values())

    def evaluate(self, point):
        acc = 0
        for exponents, coeff in self.dictionary.items():
            prod = coeff
            for index, exponent in enumerate(exponents):
                prod = prod * F.fpow(point[index], exponent) % P
            acc = (acc + prod) % P
        return acc

    def evaluate_symbolic(self, point):
        """m_polynomial.rs:128-142, literally (schoolbook products; slow, used on small cases)."""
        acc = []
        for exponents, coeff in self.dictionary.items():
            prod = [coeff]
            for index, exponent in enumerate(exponents):
                prod = PL.mul(prod, poly_pow(point[index], exponent))
            acc = PL.add(acc, prod)
        return acc

    def evaluate_symbolic_grouped(self, point, mul=None):
        """The same polynomial as evaluate_symbolic (exact arithmetic, so any order of operations
        gives the same coefficients up to trailing zeros), computed by grouping the terms by their
        exponents in the variables 1.. : sum_g (sum_terms coeff * x0^e0) * prod_i point[i]^e_i.
        `mul` multiplies two coefficient lists (default: schoolbook)."""
        mul = mul or PL.mul
        groups = {}
        for exponents, coeff in self.dictionary.items():
            e = tuple(exponents) + (0,) * (len(point) - len(exponents))
            groups.setdefault(e[1:], {})
            groups[e[1:]][e[0]] = (groups[e[1:]].get(e[0], 0) + coeff) % P
        powers = {}

        def power(i, k):
            if (i, k) not in powers:
                powers[(i, k)] = [1] if k == 0 else mul(power(i, k - 1), point[i])
            return powers[(i, k)]
        acc = []
        for rest, by_e0 in groups.items():
            # the polynomial in point[0] alone: sum coeff * point[0]^e0
            inner = []
            for e0, coeff in by_e0.items():
                inner = PL.add(inner, PL.mul([coeff], power(0, e0)))
            prod = inner
            for i, k in enumerate(rest, start=1):
                if k:
                    prod = mul(prod, power(i, k))
            acc = PL.add(acc, prod)
        return acc

    def __eq__(self, other):
        return self.dictionary == other.dictionary

    def __neg__(self):
        return MPolynomial({k: (-v) % P for k, v in self.dictionary.items()})

    def __add__(self, rhs):
        if not self.dictionary:
            return MPolynomial(rhs.dictionary)
        if not rhs.dictionary:
            return MPolynomial(self.dictionary)
        nv = max(max(len(k) for k in self.dictionary), max(len(k) for k in rhs.dictionary))
        d = {}
        for k, v in self.dictionary.items():
            d[tuple(k) + (0,) * (nv - len(k))] = v           # insert (a later duplicate key overwrites)
        for k, v in rhs.dictionary.items():
            k = tuple(k) + (0,) * (nv - len(k))
            d[k] = (d[k] + v) % P if k in d else v
        return MPolynomial(d)

    def __sub__(self, rhs):
        return self + (-rhs)

    def __mul__(self, rhs):
        nv = max(max(len(k) for k in self.dictionary), max(len(k) for k in rhs.dictionary))
        d = {}
        for k0, v0 in self.dictionary.items():
            for k1, v1 in rhs.dictionary.items():
                e = [0] * nv
                for i, v in enumerate(k0):
                    e[i] += v
                for i, v in enumerate(k1):
                    e[i] += v
                e = tuple(e)
                d[e] = (d[e] + v0 * v1) % P if e in d else v0 * v1 % P
        return MPolynomial(d)

    def __pow__(self, exponent):
        if self.is_zero():
            return MPolynomial.zero()
        nv = len(next(iter(self.dictionary)))
        acc = MPolynomial({(0,) * nv: 1})
        for b in _bits_msb_first(exponent):
            acc = acc * acc
            if b:
                acc = acc * self
        return acc


def poly_pow(p, exponent):
    """Polynomial ^ u128, polynomial.rs:329-354."""
    if PL.degree(p) is None:
        return []
    acc = [1]
    if exponent == 0:
        return acc
    for i in range(bit_count(exponent) - 1, -1, -1):
        acc = PL.mul(acc, acc)
        if (1 << i) & exponent:
            acc = PL.mul(acc, p)
    return acc


# ---- utils/matrix.rs -------------------------------------------------------------------------
def rref(matrix):
    lead = 0
    rows, cols = len(matrix), len(matrix[0])
    for r in range(rows):
        if cols <= lead:
            break
        i = r
        stop = False
        while matrix[i][lead] % P == 0:
            i += 1
            if rows == i:
                i = r
                lead += 1
                if cols == lead:
                    stop = True
                    break
        if stop:
            break
        matrix[i], matrix[r] = matrix[r], matrix[i]
        if matrix[r][lead] % P != 0:
            d = matrix[r][lead]
            matrix[r] = [F.div(el, d) for el in matrix[r]]
        for i in range(rows):
            if i != r:
                hold = matrix[i][lead]
                matrix[i] = [(matrix[i][k] - hold * matrix[r][k]) % P for k in range(cols)]
        lead += 1


def transpose(matrix):
    return [[matrix[row][col] for row in range(len(matrix))] for col in range(len(matrix[0]))]


def inverse(matrix):
    n = len(matrix)
    m = []
    for i, row in enumerate(matrix):
        assert len(row) == n, "Inverse exists only for square matrices"
        m.append(list(row) + [1 if j == i else 0 for j in range(n)])
    rref(m)
    for i, row in enumerate(m):
        assert all(el == 0 for el in row[:i]) and row[i] == 1 and all(el == 0 for el in row[i + 1:n]), \
            "Couldnt construct identity matrix to find inverse"
    return [row[n:] for row in m]
