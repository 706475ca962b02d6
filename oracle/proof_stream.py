"""Proof stream wire format + Fiat-Shamir (oracle; test infrastructure only).

Restates src/proof_stream.rs:14-83, src/stark/proof_stream_enum.rs:7-190,
src/rescue_prime/proof_stream.rs:9-62, src/utils/digest.rs:17-33 and
src/crypto/shake256.rs:7-19.

Objects are tuples ``(kind, payload)`` with kind in ROOT/CODEWORD/PATH/LEAFS/VALUE.
"""
import hashlib
from . import field as F

ROOT, CODEWORD, PATH, LEAFS, VALUE = 0, 1, 2, 3, 4
PROOF_BYTES = 32                                  # crypto/shake256.rs:5


def shake256(data: bytes, n: int) -> bytes:
    return hashlib.shake_256(data).digest(n)      # crate sha3 0.10.8 Shake256


def _be16(v):
    return int(v).to_bytes(16, "big")


def _be8(v):
    return int(v).to_bytes(8, "big")


def obj_payload(obj) -> bytes:
    # proof_stream_enum.rs:67-131 (to_bytes)
    kind, x = obj
    if kind == ROOT:
        return bytes(x)
    if kind == CODEWORD:
        return b"".join(_be16(v) for v in x)
    if kind == PATH:
        return b"".join(_be8(len(node)) + bytes(node) for node in x)
    if kind == LEAFS:
        return _be16(x[0]) + _be16(x[1]) + _be16(x[2])
    if kind == VALUE:
        return _be16(x)
    raise ValueError("Unknown code")


def digest(objects) -> bytes:
    # proof_stream_enum.rs:161-190: 16-byte BE field order if any object carries a
    # field element (Codeword/Leafs/Value), else 16 zero bytes; then per object
    # code:u8 || len:u64_be || payload.  NB an *empty* Codeword carries no field.
    has_field = False
    body = bytearray()
    for obj in objects:
        kind, x = obj
        if kind in (LEAFS, VALUE) or (kind == CODEWORD and len(x) > 0):
            has_field = True
        p = obj_payload(obj)
        body += bytes([kind]) + _be8(len(p)) + p
    return _be16(F.P if has_field else 0) + bytes(body)


def parse(data: bytes):
    """Inverse of digest() (src/stark/stark.rs:30-67 deser_independent_proof_stream)."""
    objects, pos = [], 16
    while pos < len(data):
        kind = data[pos]
        ln = int.from_bytes(data[pos + 1:pos + 9], "big")
        p = data[pos + 9:pos + 9 + ln]
        pos += 9 + ln
        if kind == ROOT:
            objects.append((ROOT, bytes(p)))
        elif kind == CODEWORD:
            objects.append((CODEWORD, [int.from_bytes(p[i:i + 16], "big") for i in range(0, ln, 16)]))
        elif kind == PATH:
            nodes, q = [], 0
            while q < ln:
                sz = int.from_bytes(p[q:q + 8], "big")
                nodes.append(bytes(p[q + 8:q + 8 + sz]))
                q += 8 + sz
            objects.append((PATH, nodes))
        elif kind == LEAFS:
            objects.append((LEAFS, tuple(int.from_bytes(p[i:i + 16], "big") for i in (0, 16, 32))))
        elif kind == VALUE:
            objects.append((VALUE, int.from_bytes(p, "big")))
        else:
            raise ValueError("Unknown code")
    return objects


class IndependentProofStream:
    # proof_stream.rs:14-83
    def __init__(self, objects=None):
        self.objects = list(objects or [])
        self.read_index = 0

    def prefix(self) -> bytes:
        return b""

    def digest(self) -> bytes:
        return digest(self.objects)

    def fiat_shamir_prover(self, num_bytes=PROOF_BYTES) -> bytes:
        return shake256(self.prefix() + digest(self.objects), num_bytes)

    def fiat_shamir_verifier(self, num_bytes=PROOF_BYTES) -> bytes:
        return shake256(self.prefix() + digest(self.objects[:self.read_index]), num_bytes)

    def push(self, obj):
        self.objects.append(obj)

    def pull(self):
        assert self.read_index < len(self.objects), "Cannot pull, queue is empty"
        obj = self.objects[self.read_index]
        self.read_index += 1
        return obj


class SignatureProofStream(IndependentProofStream):
    # rescue_prime/proof_stream.rs:9-62: SHAKE input is prefixed by
    # u64_be(64) || BLAKE2b-512(document); the stored proof (digest()) is not.
    def __init__(self, document: bytes, objects=None):
        super().__init__(objects)
        self._prefix = hashlib.blake2b(bytes(document), digest_size=64).digest()

    def prefix(self) -> bytes:
        return _be8(len(self._prefix)) + self._prefix
