"""Field scalars (src/field/field.rs) - host helpers of the C ABI, no GPU needed."""
import ctypes

from . import _lib
from .context import P, le16, from_le16

FIELD_PRIME = P


def _u8(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(bytes(b))


class Field:
    """Field::new(FIELD_PRIME).  Only the crate's one field is supported."""

    def __init__(self, order=FIELD_PRIME):
        assert order == FIELD_PRIME, "the B200 path implements p = 1 + 407*2^119 only"
        self.order = order

    def generator(self):                                   # field.rs:41-44
        out = (ctypes.c_uint8 * 16)()
        _lib.lib().zkb_field_generator(out)
        return from_le16(out)

    def primitive_nth_root(self, n):                       # field.rs:58-71
        out = (ctypes.c_uint8 * 16)()
        rc = _lib.lib().zkb_primitive_nth_root(n, out)
        assert rc == 0, "Field doesnt have nth root of unity where n > 2^119 or not power of two."
        return from_le16(out)

    def sample(self, data):                                # field.rs:87-99
        out = (ctypes.c_uint8 * 16)()
        _lib.lib().zkb_field_sample(_u8(data), len(data), out)
        return from_le16(out)

    def mul(self, a, b):
        out = (ctypes.c_uint8 * 16)()
        _lib.lib().zkb_field_mul(le16(a), le16(b), out)
        return from_le16(out)

    def inv(self, a):
        out = (ctypes.c_uint8 * 16)()
        _lib.lib().zkb_field_inv(le16(a), out)
        return from_le16(out)

    def pow(self, a, e):
        out = (ctypes.c_uint8 * 16)()
        _lib.lib().zkb_field_pow(le16(a), e, out)
        return from_le16(out)
