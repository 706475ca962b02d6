// hosthash.hpp - host-side hashes and the proof-stream object (Fiat-Shamir lives on the
// host: the transcript during FRI commit is <= ~1.5 KB).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace zkb {

void host_blake2b512(const uint8_t* msg, size_t len, uint8_t out[64]);
void host_shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t out_len);

}  // namespace zkb

// Wire format: src/stark/proof_stream_enum.rs:67-190.  The body (objects
// only, without the 16-byte order header) is kept serialised, so a challenge costs one
// SHAKE256 over the existing bytes instead of re-serialising every object as the
// reference does (proof_stream.rs:36-40).
struct zkb_ps {
    std::vector<uint8_t> prefix;     // u64_be(64) || BLAKE2b-512(document) for SignatureProofStream
    std::vector<uint8_t> body;
    bool has_field = false;          // any Codeword(non-empty) / Leafs / Value pushed
    void push(uint8_t code, const uint8_t* payload, size_t len);
    void header(uint8_t out[16]) const;
};
