// fri_tail.cuh - interface of the persistent FRI tail kernel (fri_tail.cu), used by fri.cu.
#pragma once
#include "merkle.cuh"

namespace zkb {

#define ZKB_TAIL_CTAS 128u           // co-resident CTAs (cooperative launch; one per SM, 148 SMs on B200)
#define ZKB_TAIL_THREADS 512u
#define ZKB_TAIL_MAX_LOG 17u         // first layer of at most 2^17 values: chunks of <= 1024 leaves per CTA
#define ZKB_TAIL_MAX_CHUNK 1024u
#define ZKB_TAIL_MAX_ROUNDS 18u
#define ZKB_TAIL_HOST_CW_OFF 4096u   // host_out: roots at +0 (64 B per round), last codeword at +4096
#define ZKB_TAIL_TIMEOUT_FLAG 0xDEAD0000u

struct TailArgs {
    const fe* cw_in;                 // the codeword of layer r0 - 1 (fold source), or - first_is_plain - layer r0 itself
    uint32_t first_is_plain;         // r0 == 0: the first layer is hashed as it is
    uint32_t log_n0;                 // log2 length of layer r0
    uint32_t n_rounds;               // layers r0 .. r0 + n_rounds - 1
    uint32_t r0, total_rounds;
    fe* cw[ZKB_TAIL_MAX_ROUNDS];     // codeword of layer r0 + k (cw[0] == cw_in when first_is_plain)
    uint8_t* nodes[ZKB_TAIL_MAX_ROUNDS];   // retained tree of layer r0 + k (TreeLayout with top == 0)
    DevPow winv;                     // powers of omega_0^-1
    FsDev* fs;
    uint32_t* bar;                   // grid barrier counter (zero at launch)
    uint8_t* host_out;               // mapped pinned host memory or nullptr
    volatile uint32_t* host_flag;
    uint32_t seq;
    unsigned long long* dbg;         // ZKB_TAIL_DEBUG: 8 clock stamps per round written by CTA 0 (or nullptr)
};

int fri_tail_device_init(zkb_ctx* c);
int fri_tail_launch(zkb_ctx* c, const TailArgs& a);

}  // namespace zkb
