// hosthash.hpp - host-side hashes and the proof-stream object (Fiat-Shamir lives on the
// host: the transcript during FRI commit is <= ~1.5 KB).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <new>
#include <vector>

namespace zkb {

void host_blake2b512(const uint8_t* msg, size_t len, uint8_t out[64]);
void host_shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t out_len);

}  // namespace zkb

// Wire format: src/stark/proof_stream_enum.rs:67-190.  The body (objects
// only, without the 16-byte order header) is kept serialised, so a challenge costs one
// SHAKE256 over the existing bytes instead of re-serialising every object as the
// reference does (proof_stream.rs:36-40).
namespace zkb {
// Allocator of a proof stream's body: large bodies (a proof is ~1.2 MB) live in PINNED host memory so that the device-written
// wire-format segments are copied straight to their final address (one DMA, no staging buffer, no host memcpy); small ones, and
// any allocation made without a usable CUDA device, come from malloc.  A 64-byte header in front of the block remembers which.
// resize() leaves new bytes uninitialised (they are about to be overwritten by a copy).
void* ps_body_alloc(size_t bytes);
void ps_body_free(void* p);
template <typename T>
struct PsBodyAlloc {
    typedef T value_type;
    PsBodyAlloc() = default;
    template <typename U> PsBodyAlloc(const PsBodyAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(ps_body_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t) { ps_body_free(p); }
    template <typename U> void construct(U*) noexcept {}                               // default-init: leave the bytes alone
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(static_cast<A&&>(a)...); }
    template <typename U> bool operator==(const PsBodyAlloc<U>&) const { return true; }
    template <typename U> bool operator!=(const PsBodyAlloc<U>&) const { return false; }
};
typedef std::vector<uint8_t, PsBodyAlloc<uint8_t>> PsBody;
}  // namespace zkb

struct zkb_ps;
namespace zkb {
// grow a proof stream's body by `bytes` uninitialised bytes and return their address (the target of a device-to-host copy)
uint8_t* ps_body_extend(zkb_ps* ps, size_t bytes);
}
struct zkb_ps {
    std::vector<uint8_t> prefix;     // u64_be(64) || BLAKE2b-512(document) for SignatureProofStream
    zkb::PsBody body;
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
