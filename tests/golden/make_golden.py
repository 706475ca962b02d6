#!/usr/bin/env python
"""Regenerates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

reference_kats.json : known-answer vectors copied from the reference's own inline tests
                      (file:line recorded per entry); the oracle is asserted against them in
                      tests/test_oracle_kat.py and the CUDA path in tests/test_gpu_parity.py.
oracle_vectors.json : outputs of the pinned oracle on seeded synthetic inputs (SURVEY.md 8d
                      generator) - SHA-256 of NTT / LDE outputs, Merkle roots, FRI roots and
                      proof-stream digests - so the GPU box can check the CUDA path against
                      committed values even for cases the oracle would take long to redo.
The reference itself cannot be run here (Rust crate, no rustc/cargo in the image), so
these are oracle outputs, not reference outputs; the reference's KATs pin the oracle.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cbind as C, field as F, fastfri, proof_stream as PS      # noqa: E402
from oracle.fri import FRI                                                   # noqa: E402
import test_oracle_kat as K                                                  # noqa: E402


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    kats = {
        "ntt16": {"src": "src/fft/ntt.rs:84-105", "in": [str(v) for v in K.NTT16_IN], "out": [str(v) for v in K.NTT16_OUT]},
        "intt16": {"src": "src/fft/ntt.rs:114-130", "values": [str(v) for v in K.INTT16_VALUES], "coeffs": [str(v) for v in K.INTT16_COEFFS]},
        "merkle": {"src": "src/merkle_root.rs:107-244", "commit": [
            {"leafs": [11], "root": K.H_11}, {"leafs": [5462], "root": K.H_5462},
            {"leafs": [5462, 456], "root": K.H_5462_456}, {"leafs": [652, 23409], "root": K.H_652_23409},
            {"leafs": [5462, 456, 652, 23409], "root": K.H_4}],
            "open": {"index": 1, "leafs": [5462, 456, 652, 23409], "path": [K.H_5462, K.H_652_23409]}},
        "scale": {"src": "src/field/polynomial.rs:632-652", "coeffs": [1, 2, 3], "factor": 4, "out": [1, 8, 48]},
        "sample_indices": {"src": "src/fri.rs:438-447", "seed": "d4b6e8af1114859c1c24b6496a3aef2f55a21105bc103af7e12dc3b2c101fe66",
                           "size": 128, "reduced_size": 128, "number": 17,
                           "out": [40, 121, 5, 113, 97, 68, 126, 88, 26, 82, 81, 91, 93, 125, 10, 57, 48]},
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    vec = {"generator": "x_j = ((splitmix64(s,2j) << 64) | splitmix64(s,2j+1)) mod p  (oracle.cbind.synth)", "ntt": [], "lde": [], "merkle": [], "fri": []}
    for log_n in (4, 10, 13, 16, 20, 22):
        n = 1 << log_n
        x = C.synth(0x5EED0002, n)
        w = F.primitive_nth_root(n)
        y = C.ntt(w, x)
        vec["ntt"].append({"seed": 0x5EED0002, "log_n": log_n, "sha256_out": sha(y.tobytes())})
    for log_n, n_coeffs in ((10, 256), (12, 1000), (16, 1 << 14), (20, 1 << 18), (22, 1 << 20)):
        n = 1 << log_n
        x = C.synth(0x5EED0003, n_coeffs)
        w = F.primitive_nth_root(n)
        y = C.coset_lde(w, n, F.GENERATOR, x)
        vec["lde"].append({"seed": 0x5EED0003, "log_n": log_n, "n_coeffs": n_coeffs, "sha256_out": sha(y.tobytes()),
                           "merkle_root": C.merkle(y).hex()})
    for log_n in (0, 1, 5, 10, 11, 15, 16, 20):
        n = 1 << log_n
        vec["merkle"].append({"seed": 0x5EED0004, "log_n": log_n, "root": C.merkle(C.synth(0x5EED0004, n)).hex()})
    for log_n, ncc, doc in ((8, 17, None), (12, 64, None), (12, 64, "signed document"), (16, 64, None), (20, 64, None)):
        n = 1 << log_n
        w = F.primitive_nth_root(n)
        cw = C.coset_lde(w, n, F.GENERATOR, C.synth(0x5EED0003, n // 4))
        fri = FRI(F.GENERATOR, w, n, 4, ncc)
        ps = PS.SignatureProofStream(doc.encode()) if doc else PS.IndependentProofStream()
        top, codewords, trees = fastfri.prove(fri, cw, ps)
        assert fri.verify(PS.SignatureProofStream(doc.encode(), ps.objects) if doc else PS.IndependentProofStream(ps.objects), []) is None
        d = ps.digest()
        vec["fri"].append({"seed": 0x5EED0003, "log_n": log_n, "ef": 4, "ncc": ncc, "document": doc,
                           "roots": [t.root.hex() for t in trees], "top_indices": top,
                           "proof_bytes": len(d), "sha256_proof": sha(d)})
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
