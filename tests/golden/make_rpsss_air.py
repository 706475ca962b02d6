#!/usr/bin/env python
"""Regenerates tests/golden/rpsss_air.json.  Run from the repo root: python tests/golden/make_rpsss_air.py

The Rescue-Prime AIR at the tutorial's RPSSS parameters (src/rpsss.rs:103: expansion factor 4, 64 colinearity
checks, security level 128, constraint degree 3) as plain DATA, produced by the oracle's restatement of
src/rescue_prime/rescue_prime.rs (pinned by the reference's KATs, tests/test_oracle_stark.py):
  * the two transition constraints as MPolynomial dictionaries (rescue_prime.rs:244-279),
  * for a few secret keys: the hash trace (rescue_prime.rs:185-204), the public key and the boundary conditions
    (rescue_prime.rs:281-296), the document / randomness seed used, and SHA-256 + length of the signature the
    oracle's coefficient-form prover (oracle/stark.py, restating stark.rs:276-563) produces for them.
Rescue-Prime is the AIR's author and out of scope for the CUDA path; bench.py's `signatures`
workload and tests feed these to zk_stark_tutor_b200.Stark.prove, which must reproduce the digests."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.stark import RPSSS, deterministic_rng      # noqa: E402


def main():
    r = RPSSS(4, 64, 128, 3)
    tcs = r.transition_constraints()
    out = {
        "src": "src/rpsss.rs:103 parameters; constraints rescue_prime.rs:244-279; traces rescue_prime.rs:185-204",
        "params": {"expansion_factor": 4, "num_collinearity_checks": 64, "security_level": 128, "num_registers": r.rp.m,
                   "num_cycles": r.rp.N + 1, "transition_constraints_degree": 3},
        "transition_constraints": [[[list(k), str(v)] for k, v in tc.dictionary.items()] for tc in tcs],
        "cases": [],
    }
    for i in range(4):
        sk, pk = r.keygen(deterministic_rng(b"key-%d" % i))
        doc = b"document %d" % i
        seed = b"sign-%d" % i
        sig = r.sign(sk, doc, deterministic_rng(seed))
        assert r.verify(pk, doc, sig) is None
        out["cases"].append({"secret_key": str(sk), "public_key": str(pk), "document": doc.decode(), "rng_seed": seed.decode(),
                             "trace": [[str(v) for v in row] for row in r.rp.trace(sk)],
                             "boundary": [[c, reg, str(v)] for c, reg, v in r.rp.boundary_constraints(pk)],
                             "signature_bytes": len(sig), "signature_sha256": hashlib.sha256(sig).hexdigest()})
    with open(os.path.join(HERE, "rpsss_air.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote rpsss_air.json:", sum(len(t) for t in out["transition_constraints"]), "terms,", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
