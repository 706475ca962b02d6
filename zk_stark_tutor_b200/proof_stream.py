"""Proof streams: the mirror of src/proof_stream.rs (IndependentProofStream),
src/rescue_prime/proof_stream.rs (SignatureProofStream) and the wire format of
src/stark/proof_stream_enum.rs.  Prover side only (push / digest / fiat_shamir_prover);
objects are also kept as Python tuples so tests can feed them to a verifier."""
import ctypes

from . import _lib
from .context import le16, pack

ROOT, CODEWORD, PATH, LEAFS, VALUE = 0, 1, 2, 3, 4      # proof_stream_enum.rs codes
PROOF_BYTES = 32                                        # crypto/shake256.rs:5


def _u8(b):
    return (ctypes.c_uint8 * max(len(b), 1)).from_buffer_copy(bytes(b) or b"\0")


class IndependentProofStream:
    def __init__(self, _document=None):
        self.lib = _lib.lib()
        h = ctypes.c_void_p()
        if _document is None:
            rc = self.lib.zkb_ps_create(None, 0, 0, ctypes.byref(h))
        else:
            rc = self.lib.zkb_ps_create(_u8(_document), len(_document), 1, ctypes.byref(h))
        assert rc == 0
        self.h = h
        self.objects = []

    def push(self, obj):
        kind, x = obj
        if kind == ROOT:
            self.lib.zkb_ps_push_root(self.h, _u8(x), len(x))
        elif kind == CODEWORD:
            a = pack(list(x))
            self.lib.zkb_ps_push_codeword(self.h, a.ctypes.data, len(x))
        elif kind == PATH:
            self.lib.zkb_ps_push_path(self.h, _u8(b"".join(x)), len(x))
        elif kind == LEAFS:
            self.lib.zkb_ps_push_leafs(self.h, le16(x[0]), le16(x[1]), le16(x[2]))
        elif kind == VALUE:
            self.lib.zkb_ps_push_value(self.h, le16(x))
        else:
            raise ValueError("Unknown code")
        if self.objects is not None:       # None once a C-side call (zkb_fri_prove) has appended objects itself
            self.objects.append(obj)

    def digest(self):
        n = self.lib.zkb_ps_digest(self.h, None, 0)
        buf = (ctypes.c_uint8 * max(n, 1))()
        self.lib.zkb_ps_digest(self.h, buf, n)
        return bytes(buf)[:n]

    def fiat_shamir_prover(self, num_bytes=PROOF_BYTES):
        out = (ctypes.c_uint8 * num_bytes)()
        self.lib.zkb_ps_fiat_shamir(self.h, num_bytes, out)
        return bytes(out)

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.zkb_ps_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SignatureProofStream(IndependentProofStream):
    def __init__(self, document):
        super().__init__(bytes(document))
