"""F_p arithmetic, p = 1 + 407*2^119 (oracle; test infrastructure only).

Restates the *results* of src/field/field.rs:101-169 and
src/field/field_element.rs:46-143 on canonical values (< p).  The reference's
mul_mod is a bit-serial double-and-add and its inverse an unsigned xgcd; both are
exact, so plain big-int ``%`` / ``pow(a, -1, p)`` give identical values.
"""

P = 1 + 407 * (1 << 119)                     # field.rs:9-10
GENERATOR = 85408008396924667383611388730472331217  # field.rs:43 (order exactly 2^119)
TWO_ADICITY = 119


def add(a, b):
    return (a + b) % P                        # field.rs:109-115


def sub(a, b):
    return (a - b) % P                        # field.rs:101-107


def neg(a):
    return (-a) % P                           # field.rs:133-139


def mul(a, b):
    return (a * b) % P                        # field.rs:117-131


def inv(a):
    # field.rs:160-169 (u_xgcd); inverse(0) yields 0 there - never hit on the path.
    if a % P == 0:
        return 0
    return pow(a, -1, P)


def div(a, b):
    assert b != 0, "divide by zero"           # field_element.rs:85
    return mul(a, inv(b))


def fpow(a, e):
    # field_element.rs:108-143: MSB-first square-and-multiply; a^0 = 1 (also 0^0 = 1).
    return pow(a, e, P)


def primitive_nth_root(n):
    # field.rs:58-71: repeated squaring of GENERATOR from order 2^119 down to n.
    assert n & (n - 1) == 0 and n <= (1 << 119), \
        "Field does not have any roots where n > 2^119 or not a power of two."
    root = GENERATOR
    order = 1 << 119
    while order != n:
        root = mul(root, root)
        order >>= 1
    return root


def sample(data: bytes):
    # field.rs:87-99: acc = (acc << 8 wrapping) ^ b  ==> big-endian value of the
    # last 16 bytes, then % p.
    acc = 0
    for b in data:
        acc = ((acc << 8) & ((1 << 128) - 1)) ^ b
    return acc % P


def to_le16(v):
    return int(v).to_bytes(16, "little")


def from_le16(b):
    return int.from_bytes(b, "little")


# --- the synthetic-input generator shared by oracle, C oracle and device tests ----
# x_j = ((splitmix64(s, 2j) << 64) | splitmix64(s, 2j+1)) mod p
_M64 = (1 << 64) - 1


def splitmix64(seed, idx):
    z = (seed + (idx + 1) * 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_elements(seed, n, start=0):
    return [((splitmix64(seed, 2 * j) << 64) | splitmix64(seed, 2 * j + 1)) % P
            for j in range(start, start + n)]
