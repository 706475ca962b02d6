/* zkb200.h - C ABI of libzkb200.so: the B200-native NTT / coset-LDE / Merkle / FRI engine
 * that sits behind the function signatures of SpekalsG3/zk-stark-tutor's hot path.
 *
 * The reference is a Rust crate with NO FFI of its own; its "interface" for this path is a
 * set of plain `pub fn`s.  Each entry point below names the reference function whose BODY
 * it replaces (file:line inside the reference).  INTEGRATION.md shows the Rust-side
 * `extern "C"` block and the shim bodies a maintainer would add.
 *
 * Conventions
 *  - A field element is 16 bytes, little-endian u128, canonical (< p = 1 + 407*2^119):
 *    the same bytes as Rust's in-memory `u128` on x86-64.  `FieldElement` itself is not
 *    repr(C) ({&Field, u128}); the shim packs `.value`s into a contiguous [u128].
 *  - A digest / Merkle node is 64 raw bytes (BLAKE2b-512).
 *  - Every data pointer may be a HOST pointer or a DEVICE pointer of the context's GPU;
 *    the library detects which (cudaPointerGetAttributes).  Host buffers are staged
 *    through the context's stream inside the call; device buffers are used in place, so a
 *    pipeline LDE -> commit -> FRI never leaves HBM.
 *  - All functions return 0 on success or a negative ZKB_ERR_* code; zkb_last_error()
 *    gives the message.  The reference PANICS on the same misuse; the Rust
 *    shim turns a non-zero status into panic!(message).
 *  - There is no CPU fallback: without a CUDA device zkb_ctx_create fails.
 *  - A context is used by one thread at a time (the reference is single-threaded).
 */
#ifndef ZKB200_H
#define ZKB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_OK 0
#define ZKB_ERR_CUDA (-1)        /* CUDA runtime error (message has the call)                 */
#define ZKB_ERR_ARG (-2)         /* null / inconsistent argument                              */
#define ZKB_ERR_EMPTY (-3)       /* ntt on empty input        (ntt.rs:11 index panic)         */
#define ZKB_ERR_NOT_POW2 (-4)    /* Merkle length not 2^k     (merkle_root.rs:9,36)           */
#define ZKB_ERR_TOO_LONG (-5)    /* coeffs.len() > root_order (ntt_arithmetics.rs:168)        */
#define ZKB_ERR_ROOT_ORDER (-6)  /* root^order != 1 or not primitive (ntt_arithmetics.rs:11-24) */
#define ZKB_ERR_DIV_ZERO (-7)    /* divide by zero (field_element.rs:85, ntt_arithmetics.rs:258) */
#define ZKB_ERR_INDEX (-8)       /* opening index out of range (merkle_root.rs:37)            */
#define ZKB_ERR_LENGTH (-9)      /* domain/codeword length mismatch (fri.rs:215-219)          */
#define ZKB_ERR_ROUNDS (-10)     /* FRI needs >= 2 rounds (fri.rs:225 unwrap)                 */
#define ZKB_ERR_DEGREE (-11)     /* divide by polynomial of larger degree (ntt_arithmetics.rs:268) */
#define ZKB_ERR_CALLBACK (-12)   /* Fiat-Shamir callback returned non-zero                    */
#define ZKB_ERR_NOMEM (-13)      /* host allocation failed (std::bad_alloc); a proof stream that was being appended to is unusable */

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_tree zkb_tree;
typedef struct zkb_fri_layers zkb_fri_layers;
typedef struct zkb_ps zkb_ps;
struct zkb_fri_params;

/* ---- context ------------------------------------------------------------------------ */
/* One context per GPU.  `stream` = a cudaStream_t to run on (e.g. torch's current stream; pass
 * cudaStreamLegacy = (void*)1 for the legacy default stream), or NULL to let the context create
 * its own non-blocking stream. */
int zkb_ctx_create(int device, void* stream, zkb_ctx** out);
void zkb_ctx_destroy(zkb_ctx* ctx);
const char* zkb_last_error(const zkb_ctx* ctx);
int zkb_ctx_sync(zkb_ctx* ctx);                       /* cudaStreamSynchronize             */
uint64_t zkb_ctx_launches(const zkb_ctx* ctx);        /* kernels launched so far           */
const char* zkb_version(void);
/* Pinned (page-locked) host coefficients handed to zkb_coset_lde* / zkb_lde_fri_commit* are by default
 * read in place by the first NTT pass (the PCIe transfer overlaps that pass).  A caller that keeps several
 * contexts busy on one GPU (column / proof pipelines) gets better overlap from staged copies on the copy
 * engines: enable = 0 restores the explicit H2D copy for this context. */
int zkb_ctx_zero_copy_inputs(zkb_ctx* ctx, int enable);
/* The batched entry points (zkb_fri_prove_batch, zkb_merkle_open_ps_batch) assemble the instances' proof streams on up to
 * `threads` host threads per call (default 16).  A caller that keeps several batches in flight on its own threads should lower it
 * (8 batches in flight on a 16-core host: 1-4 threads beat 16 by 16 %). */
int zkb_ctx_assembly_threads(zkb_ctx* ctx, int threads);
/* enable != 0: the batched entry points wait for the GPU by polling with short sleeps instead of cudaStreamSynchronize's
 * spinning - for callers that keep more contexts in flight than they have idle cores. */
int zkb_ctx_blocking_sync(zkb_ctx* ctx, int enable);
/* Threads per CTA (128, 256 or 512; default 512) of the persistent kernel that runs the small layers of FRI::commit (csrc/fri_tail.cu).
 * 512 takes the whole register file of 128 SMs while that latency-bound kernel runs - best for one codeword at a time.  A caller that
 * keeps several contexts busy on one GPU (column pipelines) sets 0: no persistent kernel, every round as its own small launches (the
 * Fiat-Shamir step stays on the device, nothing waits for the host), which the other contexts' kernels interleave with. */
int zkb_ctx_tail_threads(zkb_ctx* ctx, int threads);
/* Per-kernel-class device timing (CUDA events on the context's stream around every launch;
 * this is what bench.py's roofline.achieved is computed from).  enable: 0 = off, 1 = on,
 * 2 = on + reset the accumulators.  zkb_kernel_name(id) is NULL past the last class. */
int zkb_ctx_profile(zkb_ctx* ctx, int enable);
int zkb_ctx_profile_read(zkb_ctx* ctx, int kernel_id, double* total_ms, uint64_t* count);
const char* zkb_kernel_name(int kernel_id);
/* plain device-memory helpers for hosts that do not bring their own allocator */
int zkb_dev_alloc(zkb_ctx* ctx, size_t bytes, void** dptr);
int zkb_dev_free(zkb_ctx* ctx, void* dptr);
int zkb_memcpy(zkb_ctx* ctx, void* dst, const void* src, size_t bytes);   /* any direction, sync */

/* ---- field (scalar helpers, host) : src/field/field.rs:58-71,87-99,160-169 ------------ */
int zkb_primitive_nth_root(uint64_t n, uint8_t out[16]);  /* Field::primitive_nth_root     */
void zkb_field_generator(uint8_t out[16]);                /* Field::generator field.rs:41  */
void zkb_field_mul(const uint8_t a[16], const uint8_t b[16], uint8_t out[16]);
void zkb_field_inv(const uint8_t a[16], uint8_t out[16]);
void zkb_field_pow(const uint8_t a[16], uint64_t e, uint8_t out[16]);
void zkb_field_sample(const uint8_t* bytes, size_t len, uint8_t out[16]); /* Field::sample  */

/* ---- NTT : src/fft/ntt.rs ------------------------------------------------------------ */
/* ntt(root, inputs) ntt.rs:7-49.  n_in >= 1 values in; out receives next_pow2(n_in) values
 * (zero padded before the transform, natural order out).  n_in == 1 copies.  `root` must be
 * a primitive next_pow2(n_in)-th root (not checked, as in the reference). */
int zkb_ntt(zkb_ctx* ctx, const uint8_t root[16], const void* in, size_t n_in, void* out);
/* intt(root, input) ntt.rs:51-68: identity for n_in < 2, else ntt(root^-1) * n^-1. */
int zkb_intt(zkb_ctx* ctx, const uint8_t root[16], const void* in, size_t n_in, void* out);
/* `batch` independent transforms of the same length: column c starts at in + c*in_stride
 * elements / out + c*out_stride elements (trace columns). */
int zkb_ntt_batch(zkb_ctx* ctx, const uint8_t root[16], int inverse, const void* in, size_t n_in,
                  size_t in_stride, void* out, size_t out_stride, size_t batch);

/* `count` interleaved transforms of length n (power of two <= 4096): element j of sequence q is
 * in[j*stride + q]; same layout out; DEVICE pointers.  The cross-GPU stage of the four-step NTT
 * (n = number of GPUs, after the all-to-all). */
int zkb_ntt_strided(zkb_ctx* ctx, const uint8_t root[16], int inverse, const void* in, size_t n, size_t stride,
                    size_t count, void* out);

/* ---- one NTT across the GPUs of a box (BASELINE configs[4]) ----------------------------------------
 * ntt / intt (ntt.rs:7-68) of N = world * n_local values as a four-step transform: rank r holds the CYCLIC slice
 * x[r + world*m]; after the transform it holds X[k2 + n_local*k1] for k2 in [r*blk, (r+1)*blk), blk = n_local / world, laid out
 * [k1 * blk + (k2 - r*blk)].  The twiddle and the all-to-all exchange are fused into the last pass of the local transform: its
 * stores go straight into the receiving GPU's HBM over NVLink (peer access inside one process, CUDA IPC between processes).
 * world: 1, 2, 4, 8 or 16. */
typedef struct zkb_ntt4 zkb_ntt4;
/* rank `rank` of `world`, on ctx's GPU; allocates the (double-buffered) receive buffer of n_local values */
int zkb_ntt4_create(zkb_ctx* ctx, uint32_t rank, uint32_t world, size_t n_local, zkb_ntt4** plan);
void zkb_ntt4_free(zkb_ntt4* plan);
/* one process holds every rank (a context per GPU; several ranks may share a GPU): plans[r] = rank r */
int zkb_ntt4_connect_local(zkb_ntt4* const* plans, size_t world);
/* one process per GPU: export this rank's 128-byte handle, all-gather the handles over any transport (rank-major), connect */
int zkb_ntt4_export(zkb_ntt4* plan, uint8_t handle[128]);
int zkb_ntt4_connect_ipc(zkb_ntt4* plan, const uint8_t* all_handles /* world x 128 bytes */);
/* steps 1-3 (local transform, twiddle, stores into the peers' buffers), asynchronous on the context's stream.  `root` = the
 * primitive (world*n_local)-th root of the WHOLE transform; inverse != 0: intt (root^-1, N^-1 applied in finish). */
int zkb_ntt4_scatter(zkb_ntt4* plan, const uint8_t root[16], int inverse, const void* x_local);
/* step 4 (world-point transforms across the received pieces) -> out (n_local values).  Every rank's scatter of this transform
 * must have completed first: the caller orders that in the stream (zkb_ntt4_run: events; between processes any stream-ordered
 * collective on the context's stream, e.g. a one-element NCCL all-reduce, or a host barrier after zkb_ctx_sync). */
int zkb_ntt4_finish(zkb_ntt4* plan, void* out);
/* scatter on every rank, events, finish on every rank (one process).  Asynchronous for device pointers. */
int zkb_ntt4_run(zkb_ntt4* const* plans, size_t world, const uint8_t root[16], int inverse, const void* const* x_local, void* const* out);
/* the whole transform in one call: ctxs[r] runs rank r (plans are created and released inside) */
int zkb_ntt_4step(zkb_ctx* const* ctxs, size_t world, const uint8_t root[16], int inverse, const void* const* x_local, size_t n_local,
                  void* const* out);

/* ---- independent columns across the GPUs of a box (BASELINE configs[3]; stark.rs:373-381 per register) -------
 * Column i (n_coeffs coefficients; host pointer, or device pointer on the GPU of ctxs[i % n_ctx]) is extended to the FRI domain and
 * FRI-committed (zkb_lde_fri_commit_ps with a fresh IndependentProofStream) on ctxs[i % n_ctx], one host thread per context;
 * roots_out receives ncols x zkb_fri_num_rounds(p) x 64 bytes.  No field data crosses NVLink. */
int zkb_lde_commit_batch(zkb_ctx* const* ctxs, size_t n_ctx, const struct zkb_fri_params* p, const void* const* cols, size_t n_coeffs,
                         size_t ncols, uint8_t* roots_out);

/* ---- polynomial helpers : src/field/polynomial.rs, src/fft/ntt_arithmetics.rs --------- */
/* Polynomial::scale polynomial.rs:109-121: out[i] = factor^i * coeffs[i] */
int zkb_poly_scale(zkb_ctx* ctx, const uint8_t factor[16], const void* coeffs, size_t n, void* out);
/* fast_coset_evaluate ntt_arithmetics.rs:161-170 (the LDE): evaluations of the polynomial
 * on offset*<omega>, out[k] <-> offset*omega^k, `order` values out.  n_coeffs <= order. */
int zkb_coset_lde(zkb_ctx* ctx, const uint8_t omega[16], uint64_t order, const uint8_t offset[16],
                  const void* coeffs, size_t n_coeffs, void* out);
int zkb_coset_lde_batch(zkb_ctx* ctx, const uint8_t omega[16], uint64_t order, const uint8_t offset[16],
                        const void* coeffs, size_t n_coeffs, size_t in_stride, void* out,
                        size_t out_stride, size_t batch);
/* fast_multiply ntt_arithmetics.rs:5-64.  *n_out receives the coefficient count
 * (deg l + deg r + 1, or 0 if an operand is the zero polynomial); `out` must hold
 * n_lhs + n_rhs values.  HOST pointers only (small operands in the reference). */
int zkb_poly_mul(zkb_ctx* ctx, const uint8_t root[16], uint64_t root_order, const void* lhs, size_t n_lhs,
                 const void* rhs, size_t n_rhs, void* out, size_t* n_out);
/* fast_coset_divide ntt_arithmetics.rs:239-310.  `out` must hold n_lhs values. HOST pointers. */
int zkb_coset_div(zkb_ctx* ctx, const uint8_t root[16], uint64_t root_order, const uint8_t offset[16],
                  const void* lhs, size_t n_lhs, const void* rhs, size_t n_rhs, void* out, size_t* n_out);

/* ---- Merkle : src/merkle_root.rs ------------------------------------------------------ */
/* MerkleRoot::commit merkle_root.rs:21-32 for T = FieldElement: leaf = BLAKE2b-512(decimal
 * ASCII of the value), node = BLAKE2b-512(left || right); n must be a power of two. */
int zkb_merkle_commit(zkb_ctx* ctx, const void* vals, size_t n, uint8_t root[64]);
/* Same, but keeps the tree on the device so that openings do not rebuild it
 * (the reference rebuilds the whole tree on every MerkleRoot::open, merkle_root.rs:55-66).
 * The tree keeps a device copy of / reference to `vals`: if `vals` is a device pointer it
 * must stay alive until zkb_merkle_free. */
int zkb_merkle_build(zkb_ctx* ctx, const void* vals, size_t n, zkb_tree** tree);
int zkb_merkle_root(const zkb_tree* tree, uint8_t root[64]);
/* MerkleRoot::open merkle_root.rs:34-66 for k indices at once: paths_out (HOST) receives
 * k * log2(n) * 64 bytes, each path leaf-sibling first.  n must be >= 2. */
int zkb_merkle_open(zkb_tree* tree, const uint64_t* idx, size_t k, uint8_t* paths_out);
void zkb_merkle_free(zkb_tree* tree);
/* MerkleRoot::verify merkle_root.rs:69-95 (host; returns 1 = accept, 0 = reject) */
int zkb_merkle_verify(const uint8_t root[64], uint64_t index, const uint8_t* path, size_t path_len,
                      const uint8_t leaf[16]);
/* BLAKE2b-512 of arbitrary bytes on the host (crypto/blake2b512.rs:4-14) */
void zkb_blake2b512(const uint8_t* msg, size_t len, uint8_t out[64]);
/* SHAKE256 XOF on the host (crypto/shake256.rs:7-19) */
void zkb_shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t out_len);
/* The same function computed by the DEVICE sponge (one warp, csrc/keccak.cuh) that draws the Fiat-Shamir challenges
 * inside the FRI kernels; out_len <= 136.  Exists so that the KATs of crypto/shake256.rs / proof_stream.rs:129-145 can be
 * run against the device code. */
int zkb_shake256_device(zkb_ctx* ctx, const uint8_t* msg, size_t len, uint8_t* out, size_t out_len);

/* ---- FRI : src/fri.rs ------------------------------------------------------------------ */
typedef struct zkb_fri_params {
    uint8_t offset[16];            /* FRI::new fri.rs:23-38 */
    uint8_t omega[16];
    uint64_t domain_length;
    uint64_t expansion_factor;
    uint64_t num_colinearity_tests;
} zkb_fri_params;

/* FRI::num_rounds fri.rs:40-50 */
uint64_t zkb_fri_num_rounds(const zkb_fri_params* p);
/* the split-and-fold step of FRI::commit fri.rs:150-159:
 * out[i] = 2^-1((1 + alpha/x_i) cw[i] + (1 - alpha/x_i) cw[n/2+i]), x_i = offset*omega^i */
int zkb_fri_fold(zkb_ctx* ctx, const void* cw, size_t n, const uint8_t alpha[16],
                 const uint8_t offset[16], const uint8_t omega[16], void* out);
/* Fiat-Shamir hook: called once per round with that round's Merkle root, in order.
 * If want_alpha != 0 it must write the challenge alpha (field element, 16-byte LE) -
 * i.e. push Root, then Field::sample(fiat_shamir_prover(32)) as fri.rs:136-146 does.
 * On the last round want_alpha == 0 (root pushed, no challenge drawn, fri.rs:140). */
typedef int (*zkb_fs_callback)(void* user, uint32_t round, const uint8_t root[64], int want_alpha,
                               uint8_t alpha_out[16]);
/* FRI::commit fri.rs:115-172 with every layer kept on the device: per round Merkle-commit
 * the codeword, hand the root to `fs`, fold with the returned alpha (fold fused with the
 * next layer's leaf hashing).  The caller pushes Codeword(last) itself (zkb_fri_last_codeword). */
int zkb_fri_commit(zkb_ctx* ctx, const zkb_fri_params* p, const void* codeword, size_t n,
                   zkb_fs_callback fs, void* user, zkb_fri_layers** layers);
/* The Stark prover's last two steps in one call (stark.rs:500-522): coset-LDE the
 * coefficient vector to p->domain_length evaluations on p->offset * <p->omega>
 * (fast_coset_evaluate ntt_arithmetics.rs:161-170), then FRI::commit on that codeword
 * without it ever leaving HBM.  Layer 0 of `layers` is the codeword. */
int zkb_lde_fri_commit(zkb_ctx* ctx, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs,
                       zkb_fs_callback fs, void* user, zkb_fri_layers** layers);
uint64_t zkb_fri_layer_count(const zkb_fri_layers* l);
uint64_t zkb_fri_layer_len(const zkb_fri_layers* l, uint64_t round);
int zkb_fri_layer_root(const zkb_fri_layers* l, uint64_t round, uint8_t root[64]);
/* copy layer `round` (a codeword of zkb_fri_layer_len values) to `out` (host or device) */
int zkb_fri_layer_codeword(zkb_fri_layers* l, uint64_t round, void* out);
const void* zkb_fri_layer_device_ptr(const zkb_fri_layers* l, uint64_t round);
/* FRI::query fri.rs:174-208 payloads for the layer pair (round, round+1) at the ncc
 * indices idx_c: leafs_out (HOST) gets ncc * 3 * 16 bytes (cur[a], cur[b], next[c]);
 * paths_out (HOST) gets, per s: path(a) || path(b) (log2(len) nodes each) || path(c in next)
 * (log2(len)-1 nodes), 64 bytes per node. */
int zkb_fri_query(zkb_fri_layers* l, uint64_t round, const uint64_t* idx_c, size_t ncc,
                  uint8_t* leafs_out, uint8_t* paths_out);
void zkb_fri_layers_free(zkb_fri_layers* l);
/* FRI::sample_indices fri.rs:85-113 (host) */
int zkb_fri_sample_indices(const uint8_t* seed, size_t seed_len, uint64_t size, uint64_t reduced_size,
                           uint64_t number, uint64_t* out);

/* ---- proof stream : src/proof_stream.rs, src/stark/proof_stream_enum.rs ------------------ */
/* IndependentProofStream (prefix == NULL) or SignatureProofStream (prefix = the document;
 * rescue_prime/proof_stream.rs:15-22 hashes it with BLAKE2b-512). */
int zkb_ps_create(const uint8_t* document, size_t document_len, int is_signature, zkb_ps** out);
void zkb_ps_free(zkb_ps* ps);
int zkb_ps_push_root(zkb_ps* ps, const uint8_t root[64], size_t len);
int zkb_ps_push_codeword(zkb_ps* ps, const void* vals_host, size_t n);
int zkb_ps_push_path(zkb_ps* ps, const uint8_t* nodes, size_t count);       /* count x 64 bytes */
int zkb_ps_push_leafs(zkb_ps* ps, const uint8_t a[16], const uint8_t b[16], const uint8_t c[16]);
int zkb_ps_push_value(zkb_ps* ps, const uint8_t v[16]);
/* StarkProofStreamEnum::to_bytes proof_stream_enum.rs:67-127 for an object the caller serialised itself
 * (code 0..4, payload in wire order: big-endian values, length-prefixed path nodes of any size). */
int zkb_ps_push_object(zkb_ps* ps, uint8_t code, const uint8_t* payload, size_t len);
/* ProofStream::digest (the wire bytes): returns the length; copies min(len, cap) bytes */
size_t zkb_ps_digest(const zkb_ps* ps, uint8_t* out, size_t cap);
/* fiat_shamir_prover(num_bytes) proof_stream.rs:36-40 */
int zkb_ps_fiat_shamir(const zkb_ps* ps, size_t num_bytes, uint8_t* out);
/* FRI::commit (fri.rs:115-172, including the final push of Codeword(last)) against the
 * library's own proof stream; the layers stay on the device for zkb_fri_query. */
int zkb_fri_commit_ps(zkb_ctx* ctx, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps,
                      zkb_fri_layers** layers);
int zkb_lde_fri_commit_ps(zkb_ctx* ctx, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs, zkb_ps* ps,
                          zkb_fri_layers** layers);
/* FRI::prove fri.rs:210-248 end to end against a proof stream: commit, push the last
 * codeword, sample the top-level indices, push all Leafs/Path objects.  top_indices_out
 * receives num_colinearity_tests indices (the function's return value in the reference). */
int zkb_fri_prove(zkb_ctx* ctx, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps,
                  uint64_t* top_indices_out);
/* The opening loop of Stark::prove, stark.rs:546-560 (`for i in quadrupled_indices: push Value(cw[i]);
 * push Path(MerkleRoot::open(i, cw))`) for one committed codeword: k indices opened in one batch from
 * the retained tree (the reference rebuilds the tree per index), objects appended to `ps` in index order. */
int zkb_merkle_open_ps(zkb_tree* tree, const uint64_t* idx, size_t k, zkb_ps* ps);

/* ---- evaluation-form quotients and combination : src/stark/stark.rs:388-519 ---------------------------------
 * The middle of Stark::prove between the committed codewords and FRI::prove.  The reference evaluates the
 * transition constraints SYMBOLICALLY (MPolynomial::evaluate_symbolic m_polynomial.rs:128-142, schoolbook
 * polynomial products), divides by the transition zerofier (fast_coset_divide stark.rs:405-416), forms the
 * x^shift products (fast_multiply stark.rs:480, 493), the weighted sum (stark.rs:500-512) and LDEs the result
 * (stark.rs:514-519).  All of that is a polynomial identity, so the same codeword is computed here pointwise on
 * the FRI coset from the codewords that are already committed - the formula the reference's verifier applies per
 * index (stark.rs:679-769) - bit-identical, and nothing leaves HBM.  Variables of a constraint, in order:
 * x, the registers at this row, the registers at the next row (stark.rs:389-399). */
typedef struct zkb_air_desc {
    uint8_t offset[16];                  /* FRI domain: x_i = offset * omega^i                                   */
    uint8_t omega[16];
    uint64_t domain_length;
    uint64_t expansion_factor;           /* omicron = omega^expansion_factor: the next row is that many indices on */
    uint32_t num_registers;
    uint32_t num_constraints;
    const uint32_t* term_counts;         /* per constraint: number of dictionary entries                          */
    const uint8_t* coefs;                /* all terms, 16 bytes each                                              */
    const uint32_t* exps;                /* all terms, (1 + 2*num_registers) exponents each (short keys zero padded) */
    const void* const* boundary_zerofiers;      /* per register: coefficient vector (host), stark.rs:196-213     */
    const size_t* boundary_zerofier_lens;
    const void* const* boundary_interpolants;   /* per register, stark.rs:215-243                                */
    const size_t* boundary_interpolant_lens;
    const void* transition_zerofier;            /* stark.rs:186-194                                               */
    size_t transition_zerofier_len;
    const uint8_t* weights;              /* (1 + 2*num_constraints + 2*num_registers) x 16 bytes, stark.rs:447-450 */
    const uint64_t* shifts;              /* num_constraints + num_registers exponents of x, stark.rs:476, 489      */
} zkb_air_desc;
/* bq_codewords: the boundary-quotient codewords (DEVICE, register s at + s*bq_stride elements); randomizer_codeword,
 * combined_out: DEVICE, domain_length values; tq_out: NULL or DEVICE num_constraints x domain_length values that
 * receive the transition-quotient codewords (for the degree check of stark.rs:451-464). */
int zkb_air_combination(zkb_ctx* ctx, const zkb_air_desc* desc, const void* bq_codewords, size_t bq_stride,
                        const void* randomizer_codeword, void* combined_out, void* tq_out);

/* The same pipeline for batches of instances of ONE AIR, and with the boundary quotients in evaluation form as well
 * (stark.rs:331-360 on the coset: bq = (t - I) / Z_B pointwise from the trace codewords).  What depends on the AIR's shape only
 * - grouped constraint terms, zerofier codewords and their inverses - lives in a handle created once. */
typedef struct zkb_air zkb_air;
typedef struct zkb_air_shape {
    uint8_t offset[16];
    uint8_t omega[16];
    uint64_t domain_length;
    uint64_t expansion_factor;
    uint32_t num_registers;
    uint32_t num_constraints;
    const uint32_t* term_counts;
    const uint8_t* coefs;
    const uint32_t* exps;
    const void* const* boundary_zerofiers;
    const size_t* boundary_zerofier_lens;
    const void* transition_zerofier;
    size_t transition_zerofier_len;
    const uint64_t* shifts;
} zkb_air_shape;
int zkb_air_create(zkb_ctx* ctx, const zkb_air_shape* shape, zkb_air** out);
void zkb_air_free(zkb_air* air);
/* the instances' boundary interpolants (stark.rs:215-243): batch x num_registers coefficient vectors of interp_len values each
 * (zero padded), instance-major; evaluated on the coset and kept in the handle for the two calls below */
int zkb_air_set_interpolants(zkb_air* air, size_t batch, const void* interpolants, size_t interp_len);
/* trace codewords -> boundary-quotient codewords; register s of instance b at + s*stride + b*inst elements (DEVICE) */
int zkb_air_boundary_quotients(zkb_air* air, size_t batch, const void* trace_codewords, size_t trace_stride, size_t trace_inst,
                               void* bq_out, size_t bq_stride, size_t bq_inst);
/* weights: batch x (1 + 2*num_constraints + 2*num_registers) x 16 bytes (host); everything else DEVICE; tq_out may be NULL */
int zkb_air_combine(zkb_air* air, size_t batch, const uint8_t* weights, const void* bq, size_t bq_stride, size_t bq_inst,
                    const void* randomizer, size_t randomizer_inst, void* combined_out, size_t out_inst, void* tq_out, size_t tq_inst);

/* Trace interpolation + evaluation on the FRI coset for `batch` columns (stark.rs:303-326, then ntt_arithmetics.rs:161-170):
 * column c holds `length` values (host or device, at values + c*stride elements) on omicron^0 .. omicron^(length-1), a PREFIX
 * of the order-omicron_order subgroup; out (DEVICE) receives `order` evaluations on offset*<omega> per column at + c*out_stride.
 * The interpolant is iNTT(values || 0...) mod the prefix zerofier - the polynomial fast_interpolate_domain returns - computed with
 * cached per-(length, omicron_order) tables.  coeffs_out: NULL, or DEVICE batch x length coefficients. */
int zkb_trace_lde_batch(zkb_ctx* ctx, const uint8_t omicron[16], uint64_t omicron_order, uint64_t length, const uint8_t omega[16],
                        uint64_t order, const uint8_t offset[16], const void* values, size_t stride, size_t batch, void* out,
                        size_t out_stride, void* coeffs_out);
/* Polynomial::degree (polynomial.rs:46-63) of the polynomials whose values on a coset offset*<omega> are the given DEVICE
 * codewords (iNTT + scan); -1 for the zero polynomial.  Used for the degree check of stark.rs:451-464. */
int zkb_coset_degree_batch(zkb_ctx* ctx, const uint8_t omega[16], const void* codewords, size_t n, size_t stride, size_t batch,
                           int64_t* degrees_out);

/* ---- batches of small independent instances (RPSSS-shaped proofs) ---------------------
 * One proof at a 4096-point FRI domain cannot fill a GPU; independent proofs advance in lockstep instead:
 * every launch carries all instances.  Per instance the results are byte-identical to the single-instance
 * entry points above.  Limits: device pointers, n <= 2^17, batch <= 4096. */
/* MerkleRoot::commit (stark.rs:373-381, 431-436) over `batch` codewords, instance b at vals + b*stride
 * elements; trees[b] are retained trees sharing one arena - tree 0 owns it: free it LAST. */
int zkb_merkle_build_batch(zkb_ctx* ctx, const void* vals, size_t n, size_t stride, size_t batch, zkb_tree** trees,
                           zkb_ps* const* ps /* NULL, or per tree a stream (or NULL) that receives Root(root), in tree order */);
/* FRI::prove (fri.rs:210-248) for `batch` codewords (instance b at codewords + b*stride elements) against
 * `batch` proof streams; top_indices_out receives batch x num_colinearity_tests indices. */
int zkb_fri_prove_batch(zkb_ctx* ctx, const zkb_fri_params* p, const void* codewords, size_t n, size_t stride, size_t batch,
                        zkb_ps* const* ps, uint64_t* top_indices_out);
/* zkb_merkle_open_ps for `count` trees of one zkb_merkle_build_batch (in batch order): tree i is opened at
 * idx[i*k .. i*k+k) and appends to ps[i]; trees that share a proof stream append in increasing tree order. */
int zkb_merkle_open_ps_batch(zkb_tree* const* trees, size_t count, const uint64_t* idx, size_t k, zkb_ps* const* ps);

/* Stark::prove (stark.rs:276-563) for `batch` instances of one AIR as ONE call: randomized trace -> interpolation + LDE
 * (zkb_trace_lde_batch) -> boundary quotients (zkb_air_boundary_quotients) -> commits (zkb_merkle_build_batch) -> randomizer
 * polynomial -> weights from each transcript -> transition quotients + combination (zkb_air_combine, degree check of
 * stark.rs:451-464) -> FRI::prove (zkb_fri_prove_batch) -> Value / Path openings at the quadrupled indices
 * (zkb_merkle_open_ps_batch).  The proofs are left in `ps` (read them with zkb_ps_digest); per instance the bytes are those of
 * the call sequence above, which zk_stark_tutor_b200/stark.py `prove_batch` issues stage by stage. */
typedef struct zkb_stark_shape {
    uint8_t omicron[16];                 /* generator of the trace domain (stark.rs:74-77)                               */
    uint64_t omicron_order;
    uint64_t trace_length;               /* rows of the original trace                                                   */
    uint64_t num_randomizers;            /* random rows appended to every register column (stark.rs:286-301)             */
    uint64_t rnd_poly_len;               /* coefficients of the randomizer polynomial (stark.rs:424-432)                 */
    uint32_t num_registers;
    uint32_t num_constraints;
    const int64_t* tq_degree_bounds;     /* expected degree of every transition quotient; NULL: no degree check          */
    uint32_t num_boundary;               /* boundary conditions per instance                                             */
    const uint32_t* boundary_register;   /* register of boundary condition j                                             */
    const void* lagrange;                /* per register s (m_s conditions, in order): m_s x m_s values, row i = coefficients
                                            (low first) of the Lagrange basis polynomial of its i-th point; concatenated  */
    zkb_fri_params fri;                  /* the FRI instance (offset = the coset generator, omega, domain, ef, tests)     */
    uint64_t proof_bytes;                /* Fiat-Shamir bytes behind the weights (32, stark.rs:447)                       */
} zkb_stark_shape;
/* traces: batch x trace_length x num_registers values (host; row r, register s of instance b at (b*trace_length + r)*num_registers + s);
 * boundary_values: batch x num_boundary values (host); randomness: NULL (OS entropy, getrandom) or batch x
 * (num_randomizers*num_registers + rnd_poly_len) field elements (host): the randomizer rows (row by row, register by register), then the
 * randomizer polynomial; proof_len_out: NULL or batch lengths.  A failed degree check returns ZKB_ERR_DEGREE. */
int zkb_stark_prove_batch(zkb_ctx* ctx, zkb_air* air, const zkb_stark_shape* shape, size_t batch, const void* traces,
                          const void* boundary_values, const void* randomness, zkb_ps* const* ps, uint64_t* proof_len_out);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
