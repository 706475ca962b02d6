"""BLAKE2b-512 Merkle tree over field elements (oracle; test infrastructure only).

Restates src/merkle_root.rs:7-95 and src/crypto/blake2b512.rs:4-14.  The leaf
preimage of a FieldElement is the ASCII decimal string of its value
(src/field/field_element.rs:46-50,101-105); a node is BLAKE2b-512(left || right).
"""
import hashlib


def blake2b512(data: bytes) -> bytes:
    # crypto/blake2b512.rs:4-14 - unkeyed BLAKE2b, 64-byte digest (crate blake2 0.10.6)
    return hashlib.blake2b(data, digest_size=64).digest()


def leaf_bytes(value: int) -> bytes:
    return str(int(value)).encode("ascii")      # field_element.rs:46-50


def leaf_hash(value: int) -> bytes:
    return blake2b512(leaf_bytes(value))        # merkle_root.rs:25-30


def _levels(values):
    n = len(values)
    assert n >= 1 and n & (n - 1) == 0, "Leafs len must be power of two"   # merkle_root.rs:9
    lv = [[leaf_hash(v) for v in values]]
    while len(lv[-1]) > 1:
        p = lv[-1]
        lv.append([blake2b512(p[2 * i] + p[2 * i + 1]) for i in range(len(p) // 2)])
    return lv


def tree_levels(values):
    """All levels bottom-up: levels[0] = leaf hashes, levels[-1] = [root]."""
    return _levels(values)


def commit(values) -> bytes:
    # merkle_root.rs:7-32: recursive halves == bottom-up pairing of adjacent nodes.
    return _levels(values)[-1][0]


def open_(index, values):
    # merkle_root.rs:34-66: [sibling leaf hash, sibling subtree root at level 1, ...]
    lv = _levels(values)
    assert len(values) >= 2, "open on a 1-leaf tree recurses forever in the reference"
    assert index < len(values), "cannot open invalid index"
    path = []
    for level in lv[:-1]:
        path.append(level[index ^ 1])
        index >>= 1
    return path


def verify(root: bytes, index: int, path, value: int) -> bool:
    # merkle_root.rs:69-95
    assert index < (1 << len(path)), "Cannot verify invalid index"
    h = leaf_hash(value)
    for sib in path:
        h = blake2b512(h + sib) if index % 2 == 0 else blake2b512(sib + h)
        index >>= 1
    return h == root
