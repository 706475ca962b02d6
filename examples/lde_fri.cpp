// examples/lde_fri.cpp - the reference crate's hot path from C++ (include/zk_impl.hpp over libzkb200.so):
// coset low-degree extension of a polynomial, Merkle commitment, FRI proof, verification.
//
//   g++ -O2 -std=c++17 -Iinclude examples/lde_fri.cpp -Lzk_stark_tutor_b200/lib -lzkb200 -Wl,-rpath,$PWD/zk_stark_tutor_b200/lib -o lde_fri
//   ./lde_fri [log2 of the codeword length, default 12]
//
// The same calls in the reference (Rust): fast_coset_evaluate (src/fft/ntt_arithmetics.rs:161), MerkleRoot::commit / open / verify
// (src/merkle_root.rs:21-95), FRI::new / prove / verify (src/fri.rs:23-416) over an IndependentProofStream (src/proof_stream.rs).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "zk_impl.hpp"

using namespace zk_impl;

int main(int argc, char** argv) {
    const int log_n = argc > 1 ? std::atoi(argv[1]) : 12;
    const size_t n = (size_t)1 << log_n, expansion_factor = 4, num_colinearity_tests = 64;
    try {
        Field field(FIELD_PRIME);
        FieldElement omega = field.primitive_nth_root(n), offset = field.generator();

        std::vector<FieldElement> coefficients;                       // a polynomial of degree < n / expansion_factor
        u128 x = 0x5EED;
        for (size_t i = 0; i < n / expansion_factor; i++) {
            x = x * 6364136223846793005ULL + 1442695040888963407ULL;
            coefficients.emplace_back(&field, x % FIELD_PRIME);
        }
        auto t0 = std::chrono::steady_clock::now();
        std::vector<FieldElement> codeword = fast_coset_evaluate(omega, n, offset, Polynomial(coefficients));   // the LDE
        Bytes root = MerkleRoot::commit(codeword);
        std::vector<Bytes> path = MerkleRoot::open(5, codeword);
        if (!MerkleRoot::verify(root, 5, path, codeword[5])) { std::printf("opening does not verify\n"); return 1; }

        FRI fri(offset, omega, n, expansion_factor, num_colinearity_tests);
        IndependentProofStream proof_stream;
        std::vector<size_t> indices = fri.prove(codeword, proof_stream);
        Bytes proof = proof_stream.digest();
        auto t1 = std::chrono::steady_clock::now();

        IndependentProofStream verifier_stream(deserialize_proof(proof, &field));
        std::vector<std::pair<size_t, FieldElement>> points;
        Result ok = fri.verify(verifier_stream, points);
        std::printf("codeword 2^%d: root %s..., %zu FRI rounds, proof %zu bytes, first index %zu, prove %.2f ms, verify: %s\n", log_n,
                    root.to_hex().substr(0, 16).c_str(), fri.num_rounds(), proof.buf.size(), indices[0],
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), ok.is_ok() ? "ok" : ok.err->c_str());
        return ok.is_ok() ? 0 : 1;
    } catch (const Panic& p) {
        std::printf("panic: %s (code %d)\n", p.what(), p.code);
        return 2;
    }
}
