"""CPU oracle for the NTT / LDE / Merkle / FRI hot path of SpekalsG3/zk-stark-tutor.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import or execute it, and only as the checker (or
as the thing timed for the CPU baseline), never as a fallback for the CUDA path.

It is a *restatement* of the reference's semantics (the reference is a Rust crate
and there is no rustc/cargo in this image, so it cannot be compiled or imported):

* ``oracle.field``        <- src/field/field.rs, src/field/field_element.rs
* ``oracle.ntt``          <- src/utils/bit_reverse_copy.rs, src/fft/ntt.rs,
                             src/field/polynomial.rs:46-63,109-121,
                             src/fft/ntt_arithmetics.rs:5-64,161-170,239-310
* ``oracle.merkle``       <- src/merkle_root.rs, src/crypto/blake2b512.rs
* ``oracle.proof_stream`` <- src/proof_stream.rs, src/stark/proof_stream_enum.rs,
                             src/rescue_prime/proof_stream.rs, src/utils/digest.rs,
                             src/crypto/shake256.rs
* ``oracle.fri``          <- src/fri.rs
* ``oracle.mpoly``        <- src/m_polynomial.rs, src/utils/matrix.rs, src/utils/bit_iter.rs
* ``oracle.rescue_prime`` <- src/rescue_prime/rescue_prime.rs
* ``oracle.stark``        <- src/stark/stark.rs (prove AND verify), src/rpsss.rs: the CALLERS of the
                             hot path, restated so the real prover can be run with either backend
* ``oracle/zkoracle.c``   <- the same semantics in C (unsigned __int128) for sizes
                             where Python is too slow, plus a faithful-algorithm
                             (bit-serial mul_mod, per-element pow / xgcd) variant
                             used only as the timed CPU baseline.

Pinning: every known-answer test the reference holds for this path
(src/fft/ntt.rs:78-130, src/merkle_root.rs:107-244, src/fri.rs:426-448,
src/field/field.rs:186-241, src/field/field_element.rs:151-299,
src/crypto/blake2b512.rs:22-30, src/proof_stream.rs:129-145, the 1,156,888-byte
proof size of src/rpsss.rs:89) is asserted in tests/test_oracle_kat.py; the callers' known
answers (src/rescue_prime/rescue_prime.rs:298-423, src/m_polynomial.rs:330-560,
src/utils/matrix.rs:110-183, the sign/verify round trip of src/rpsss.rs:103-135) in
tests/test_oracle_stark.py.
Third-party hashes the reference takes from crates.io (blake2 0.10.6 Blake2b512,
sha3 0.10.8 Shake256, both pinned in Cargo.lock) are the standard RFC 7693 /
FIPS 202 functions; here they come from ``hashlib`` and are pinned by the same KATs.
Whole FRI transcripts / STARK proofs have no golden vector in the reference (its
prover draws from thread_rng): for those, parity is "unpinned by the reference's
own tests" and rests on this oracle + component KATs + verifier acceptance.
"""
