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
    // Incremental SHAKE256 sponge over the transcript prefix || header || body (append-only except for the one
    // header flip when the first field element is pushed, which restarts it): a challenge absorbs only the bytes
    // added since the previous one, where the reference re-hashes the whole stream (proof_stream.rs:36-40).
    uint64_t sp_st[25] = {0};
    size_t sp_absorbed = 0;          // transcript bytes absorbed so far (a multiple of the 136-byte rate)
    bool sp_field = false;           // has_field at the time the sponge was started
    void sponge_sync();              // absorb every complete block of the current transcript
    size_t transcript_len() const { return prefix.size() + 16 + body.size(); }
    void transcript_read(size_t off, size_t len, uint8_t* out) const;
};

namespace zkb {
struct FsSponge;
// Sponge state + partial block of `ps`'s current transcript, for the device Fiat-Shamir (keccak.cuh)
void ps_export_sponge(zkb_ps* ps, FsSponge* out);
}
