// zk_impl.hpp - C++ host side of the drop-in: the public API of the reference crate `zk_impl`
// (SpekalsG3/zk-stark-tutor) for the hot path, with every body a call into libzkb200.so
// (include/zkb200.h).  Header-only, C++17, no CUDA headers needed by the includer.
//
// Why it exists: the reference is compiled Rust and the build image has no rustc / cargo, so the
// host side above the C ABI is written in C++ with the crate's own names, argument order and
// failure behaviour (INTEGRATION.md holds the Rust `extern "C"` shim a maintainer would add; this
// header is the same shim, compiled and tested).  A reference `panic!` becomes a `zk_impl::Panic`
// exception carrying the reference's message; `Result<_, String>` becomes `Result`.
//
//   reference item                                         file:line                     here
//   Field / FIELD_PRIME / generator / primitive_nth_root   src/field/field.rs:9-99       Field
//   FieldElement + - * / neg ^ inverse, Into<Bytes>        src/field/field_element.rs    FieldElement
//   Polynomial (degree, evaluate, scale, + - * %, ...)     src/field/polynomial.rs       Polynomial
//   ntt / intt                                             src/fft/ntt.rs:7-68           ntt, intt
//   fast_multiply / fast_zerofier / fast_evaluate_domain / src/fft/ntt_arithmetics.rs    same names
//   fast_coset_evaluate / fast_interpolate_domain / fast_coset_divide
//   MerkleRoot::commit / open / verify                     src/merkle_root.rs:21-95      MerkleRoot
//   ProofStream, IndependentProofStream                    src/proof_stream.rs           same names
//   SignatureProofStream                                   src/rescue_prime/proof_stream.rs
//   StarkProofStreamEnum (+ wire format)                   src/stark/proof_stream_enum.rs
//   FRI::new / num_rounds / evaluate_domain / prove / verify   src/fri.rs:23-416         FRI
//
// The verifier (FRI::verify) is scalar host code, as in the reference; only its Merkle commit of the
// last codeword and its iNTT / NTT run on the GPU.  There is no CPU fallback for the GPU-backed
// functions: without a CUDA device the first one of them throws.
#ifndef ZK_IMPL_HPP
#define ZK_IMPL_HPP
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkb200.h"

namespace zk_impl {

using u128 = unsigned __int128;

// A reference panic!/assert!/unwrap failure.  code = the ZKB_ERR_* status when it came from the library.
struct Panic : std::runtime_error {
    int code;
    explicit Panic(const std::string& m, int c = 0) : std::runtime_error(m), code(c) {}
};

// ---- utils/bytes.rs ---------------------------------------------------------------------------
struct Bytes {
    std::vector<uint8_t> buf;
    Bytes() = default;
    explicit Bytes(std::vector<uint8_t> b) : buf(std::move(b)) {}
    Bytes(const uint8_t* p, size_t n) : buf(p, p + n) {}
    // `"7aa7...".into()` in the reference's tests: a hex string (bytes.rs:26-38, 54-61)
    Bytes(const char* hex) : Bytes(unhex(hex)) {}
    Bytes(const std::string& hex) : Bytes(unhex(hex)) {}
    static Bytes zeroed(size_t n) { Bytes b; b.buf.assign(n, 0); return b; }
    const std::vector<uint8_t>& bytes() const { return buf; }
    std::string to_hex() const {
        static const char* d = "0123456789abcdef";
        std::string s;
        s.reserve(2 * buf.size());
        for (uint8_t b : buf) { s.push_back(d[b >> 4]); s.push_back(d[b & 15]); }
        return s;
    }
    Bytes operator+(const Bytes& o) const {
        Bytes r(buf);
        r.buf.insert(r.buf.end(), o.buf.begin(), o.buf.end());
        return r;
    }
    bool operator==(const Bytes& o) const { return buf == o.buf; }
    bool operator!=(const Bytes& o) const { return buf != o.buf; }

private:
    static std::vector<uint8_t> unhex(const std::string& s) {
        auto nib = [](char c) -> int {
            if (c >= '0' && c <= '9') return c - '0';
            if (c >= 'a' && c <= 'f') return c - 'a' + 10;
            if (c >= 'A' && c <= 'F') return c - 'A' + 10;
            throw Panic("invalid hex digit");
        };
        if (s.size() % 2) throw Panic("odd number of hex digits");
        std::vector<uint8_t> out(s.size() / 2);
        for (size_t i = 0; i < out.size(); i++) out[i] = (uint8_t)(nib(s[2 * i]) << 4 | nib(s[2 * i + 1]));
        return out;
    }
};

inline Bytes blake2b512(const Bytes& b) {                       // crypto/blake2b512.rs:4-14
    Bytes out = Bytes::zeroed(64);
    zkb_blake2b512(b.buf.data(), b.buf.size(), out.buf.data());
    return out;
}
constexpr size_t PROOF_BYTES = 32;                              // crypto/shake256.rs:5
inline Bytes shake256(const Bytes& b, size_t num_bytes) {       // crypto/shake256.rs:7-19
    Bytes out = Bytes::zeroed(num_bytes);
    zkb_shake256(b.buf.data(), b.buf.size(), out.buf.data(), num_bytes);
    return out;
}

// ---- the device context behind the free functions -----------------------------------------------
// The reference's functions take no context; the mirror keeps one per thread, created on first use
// (device = $ZKB_DEVICE or 0).  Hosts that manage several GPUs call zkb_* with their own contexts.
class Device {
public:
    static zkb_ctx* ctx() {
        thread_local Device d;
        return d.h_;
    }
    static void check(int rc) {
        if (rc != ZKB_OK) throw Panic(zkb_last_error(ctx()), rc);
    }

private:
    zkb_ctx* h_ = nullptr;
    Device() {
        const char* e = std::getenv("ZKB_DEVICE");
        int rc = zkb_ctx_create(e ? std::atoi(e) : 0, nullptr, &h_);
        if (rc != ZKB_OK || !h_) throw Panic(std::string("zkb_ctx_create failed (no CPU fallback): ") + zkb_last_error(nullptr), rc);
    }
    ~Device() { if (h_) zkb_ctx_destroy(h_); }
    Device(const Device&) = delete;
};

// ---- field/field.rs, field/field_element.rs ---------------------------------------------------
constexpr u128 FIELD_PRIME = ((u128)0xCB80000000000000ULL << 64) | 1ULL;   // 1 + 407 * 2^119, field.rs:9-10

inline void to_le16(u128 v, uint8_t out[16]) { std::memcpy(out, &v, 16); }  // little-endian hosts (x86-64, aarch64)
inline u128 from_le16(const uint8_t in[16]) { u128 v; std::memcpy(&v, in, 16); return v; }
inline u128 parse_u128(const char* dec) {
    u128 v = 0;
    for (const char* c = dec; *c; c++) {
        if (*c == '_') continue;
        if (*c < '0' || *c > '9') throw Panic("invalid decimal digit");
        v = v * 10 + (u128)(*c - '0');
    }
    return v;
}
inline std::string to_string(u128 v) {
    if (v == 0) return "0";
    char tmp[40];
    int n = 0;
    while (v) { tmp[n++] = (char)('0' + (int)(v % 10)); v /= 10; }
    std::string s;
    while (n) s.push_back(tmp[--n]);
    return s;
}

struct Field;
struct FieldElement {
    const Field* field = nullptr;
    u128 value = 0;
    FieldElement() = default;
    FieldElement(const Field* f, u128 v) : field(f), value(v) {}          // FieldElement::new: does not reduce
    static FieldElement new_(const Field& f, u128 v) { return FieldElement(&f, v); }
    bool is_zero() const { return value == 0; }
    FieldElement inverse() const;
    FieldElement operator+(const FieldElement& o) const;
    FieldElement operator-(const FieldElement& o) const;
    FieldElement operator*(const FieldElement& o) const;
    FieldElement operator/(const FieldElement& o) const;
    FieldElement operator-() const;
    FieldElement operator^(u128 exponent) const;                          // pow, field_element.rs:108-143
    bool operator==(const FieldElement& o) const;
    bool operator!=(const FieldElement& o) const { return !(*this == o); }
    operator Bytes() const {                                              // Into<Bytes>: decimal ASCII, field_element.rs:46-50
        std::string s = to_string(value);
        return Bytes(reinterpret_cast<const uint8_t*>(s.data()), s.size());
    }
};

struct Field {
    u128 order;
    // The reference tolerates any `order`; roots of unity - and the GPU path - exist for FIELD_PRIME
    // only (field.rs:42), so any other order is rejected where it would reach the library.
    explicit Field(u128 order_ = FIELD_PRIME) : order(order_) {}
    static Field new_(u128 order_) { return Field(order_); }
    void require_prime() const {
        if (order != FIELD_PRIME) throw Panic("the B200 path supports Field::new(FIELD_PRIME) only");
    }
    FieldElement zero() const { return FieldElement(this, 0); }
    FieldElement one() const { return FieldElement(this, 1); }
    FieldElement generator() const {                                      // field.rs:41-44
        if (order != FIELD_PRIME) throw Panic("Do not know generator for other fields beyond 1+407*2^119");
        uint8_t g[16];
        zkb_field_generator(g);
        return FieldElement(this, from_le16(g));
    }
    FieldElement primitive_nth_root(u128 n) const {                       // field.rs:58-71
        if (order != FIELD_PRIME) throw Panic("Unknown field, can't return root of unity");
        if (n == 0 || (n & (n - 1)) != 0 || n > ((u128)1 << 119))
            throw Panic("Field does not have nth root of unity where n > 2^119 or not power of two.");
        uint8_t r[16];
        if (zkb_primitive_nth_root((uint64_t)n, r) != ZKB_OK)
            throw Panic("Field does not have nth root of unity where n > 2^119 or not power of two.");
        return FieldElement(this, from_le16(r));
    }
    FieldElement sample(const Bytes& bytes) const {                       // field.rs:87-99
        require_prime();
        uint8_t o[16];
        zkb_field_sample(bytes.buf.data(), bytes.buf.size(), o);
        return FieldElement(this, from_le16(o));
    }
    bool operator==(const Field& o) const { return order == o.order; }
};

inline bool FieldElement::operator==(const FieldElement& o) const {
    // derive(PartialEq): compares the Field (by value) and the value
    return value == o.value && (field == o.field || (field && o.field && *field == *o.field));
}
inline FieldElement FieldElement::operator+(const FieldElement& o) const {      // field.rs:109-115
    const u128 p = field->order;
    const u128 nb = o.value == 0 ? 0 : p - o.value;                             // a + b = a - (p - b)
    return FieldElement(field, value >= nb ? value - nb : value + o.value);
}
inline FieldElement FieldElement::operator-(const FieldElement& o) const {      // field.rs:101-107
    const u128 p = field->order;
    return FieldElement(field, value >= o.value ? value - o.value : p - (o.value - value));
}
inline FieldElement FieldElement::operator-() const { return FieldElement(field, value == 0 ? 0 : field->order - value); }
inline FieldElement FieldElement::operator*(const FieldElement& o) const {      // field.rs:117-131
    field->require_prime();
    uint8_t a[16], b[16], r[16];
    to_le16(value, a);
    to_le16(o.value, b);
    zkb_field_mul(a, b, r);
    return FieldElement(field, from_le16(r));
}
inline FieldElement FieldElement::inverse() const {                             // field_element.rs:34-39, field.rs:160-169
    field->require_prime();
    uint8_t a[16], r[16];
    to_le16(value, a);
    zkb_field_inv(a, r);
    return FieldElement(field, from_le16(r));
}
inline FieldElement FieldElement::operator/(const FieldElement& o) const {      // field_element.rs:82-91
    if (o.is_zero()) throw Panic("divide by zero", ZKB_ERR_DIV_ZERO);
    return *this * o.inverse();
}
inline FieldElement FieldElement::operator^(u128 e) const {
    FieldElement acc(field, 1);
    if (e == 0) return acc;
    int top = 127;
    while (!((e >> top) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        acc = acc * acc;
        if ((e >> i) & 1) acc = acc * *this;
    }
    return acc;
}

// contiguous [u128] <-> Vec<FieldElement> (FieldElement is not repr(C) in the reference either)
inline std::vector<u128> pack_values(const std::vector<FieldElement>& v) {
    std::vector<u128> out(v.size());
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i].value;
    return out;
}
inline std::vector<FieldElement> attach_field(const Field* f, const std::vector<u128>& v, size_t n) {
    std::vector<FieldElement> out(n);
    for (size_t i = 0; i < n; i++) out[i] = FieldElement(f, v[i]);
    return out;
}
inline const Field* field_of(const std::vector<FieldElement>& v, const Field* fallback = nullptr) {
    const Field* f = v.empty() ? fallback : v[0].field;
    if (f) f->require_prime();
    return f;
}

// ---- field/polynomial.rs ------------------------------------------------------------------------
struct Polynomial {
    std::vector<FieldElement> coefficients;
    Polynomial() = default;
    explicit Polynomial(std::vector<FieldElement> c) : coefficients(std::move(c)) {}
    static Polynomial zero() { return Polynomial(); }
    std::optional<size_t> degree() const {                              // polynomial.rs:46-63
        std::optional<size_t> d;
        for (size_t i = 0; i < coefficients.size(); i++)
            if (coefficients[i].value != 0) d = i;
        return d;
    }
    bool is_zero() const { return !degree().has_value(); }
    const FieldElement* leading_coefficient() const {                   // polynomial.rs:69-74
        auto d = degree();
        if (d) return &coefficients[*d];
        return coefficients.empty() ? nullptr : &coefficients.back();
    }
    FieldElement evaluate(const FieldElement& point) const {            // polynomial.rs:76-100
        FieldElement value = point.field->zero(), xi = point.field->one();
        for (const FieldElement& c : coefficients) {
            value = value + c * xi;
            xi = xi * point;
        }
        return value;
    }
    std::vector<FieldElement> evaluate_domain(const std::vector<FieldElement>& domain) const {
        std::vector<FieldElement> out;
        out.reserve(domain.size());
        for (const FieldElement& x : domain) out.push_back(evaluate(x));
        return out;
    }
    // coef_i * factor^i (polynomial.rs:109-121); runs on the GPU
    Polynomial scale(const FieldElement& factor) const {
        if (coefficients.empty()) return Polynomial();
        const Field* f = field_of(coefficients);
        std::vector<u128> in = pack_values(coefficients), out(in.size());
        uint8_t fa[16];
        to_le16(factor.value, fa);
        Device::check(zkb_poly_scale(Device::ctx(), fa, in.data(), in.size(), out.data()));
        return Polynomial(attach_field(f, out, out.size()));
    }
    bool operator==(const Polynomial& o) const { return coefficients == o.coefficients; }   // derive(PartialEq): trailing zeros count
    bool operator!=(const Polynomial& o) const { return !(*this == o); }
    Polynomial operator-() const {
        Polynomial r(coefficients);
        for (auto& c : r.coefficients) c = -c;
        return r;
    }
    Polynomial operator+(const Polynomial& rhs) const {                 // polynomial.rs:252-281
        if (is_zero()) return rhs;
        if (rhs.is_zero()) return *this;
        const Field* f = coefficients[0].field;
        std::vector<FieldElement> out(std::max(coefficients.size(), rhs.coefficients.size()), f->zero());
        for (size_t i = 0; i < coefficients.size(); i++) out[i] = out[i] + coefficients[i];
        for (size_t i = 0; i < rhs.coefficients.size(); i++) out[i] = out[i] + rhs.coefficients[i];
        return Polynomial(std::move(out));
    }
    Polynomial operator-(const Polynomial& rhs) const { return *this + (-rhs); }
    Polynomial operator*(const Polynomial& rhs) const {                 // schoolbook, polynomial.rs:290-314
        if (coefficients.empty() || rhs.coefficients.empty()) return Polynomial();
        const Field* f = coefficients[0].field;
        std::vector<FieldElement> out(coefficients.size() + rhs.coefficients.size() - 1, f->zero());
        for (size_t i = 0; i < coefficients.size(); i++) {
            if (coefficients[i].is_zero()) continue;
            for (size_t j = 0; j < rhs.coefficients.size(); j++) out[i + j] = out[i + j] + coefficients[i] * rhs.coefficients[j];
        }
        return Polynomial(std::move(out));
    }
    // polynomial.rs:179-224: Err("Denominator is zero or empty") becomes a Panic, as `%` unwraps it
    static std::pair<Polynomial, Polynomial> divide_with_rem(const Polynomial& numerator, const Polynomial& denominator) {
        auto dd = denominator.degree();
        if (!dd) throw Panic("Denominator is zero or empty");
        auto nd = numerator.degree();
        if (!nd || *nd < *dd) return {Polynomial(), numerator};
        const Field* f = denominator.coefficients[0].field;
        Polynomial rem = numerator;
        size_t steps = *nd - *dd + 1;
        std::vector<FieldElement> q(steps, f->zero());
        const FieldElement lead = *denominator.leading_coefficient();
        for (size_t s = 0; s < steps; s++) {
            auto rd = rem.degree();
            if (!rd || *rd < *dd) break;
            FieldElement coef = *rem.leading_coefficient() / lead;
            size_t shift = *rd - *dd;
            std::vector<FieldElement> sub(shift, f->zero());
            sub.push_back(coef);
            rem = rem - Polynomial(std::move(sub)) * denominator;
            q[shift] = coef;
        }
        return {Polynomial(std::move(q)), rem};
    }
    Polynomial operator%(const Polynomial& rhs) const { return divide_with_rem(*this, rhs).second; }
    // polynomial.rs:123-148 (Lagrange, schoolbook): used by test_colinearity only
    static Polynomial interpolate_domain(const std::vector<FieldElement>& domain, const std::vector<FieldElement>& values) {
        if (domain.size() != values.size()) throw Panic("number of elements in domain does not match number of values");
        if (domain.empty()) throw Panic("Cannot interpolate between zero points");
        const Field* f = domain[0].field;
        Polynomial x({f->zero(), f->one()}), acc;
        for (size_t i = 0; i < domain.size(); i++) {
            Polynomial prod({values[i]});
            for (size_t j = 0; j < domain.size(); j++) {
                if (i == j) continue;
                prod = prod * (x - Polynomial({domain[j]})) * Polynomial({(domain[i] - domain[j]).inverse()});
            }
            acc = acc + prod;
        }
        return acc;
    }
    static bool test_colinearity(const std::vector<std::pair<FieldElement, FieldElement>>& points) {   // polynomial.rs:161-177
        std::vector<FieldElement> d, v;
        for (auto& p : points) { d.push_back(p.first); v.push_back(p.second); }
        auto deg = interpolate_domain(d, v).degree();
        return deg && *deg == 1;
    }
};

// ---- fft/ntt.rs -----------------------------------------------------------------------------------
inline size_t next_power_of_two(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }

// ntt(root, inputs) ntt.rs:7-49: zero-pads to the next power of two, natural order in and out
inline std::vector<FieldElement> ntt(const FieldElement& root, const std::vector<FieldElement>& inputs) {
    if (inputs.empty()) throw Panic("index out of bounds: the len is 0 but the index is 0", ZKB_ERR_EMPTY);    // ntt.rs:11
    const Field* f = field_of(inputs, root.field);
    std::vector<u128> in = pack_values(inputs), out(next_power_of_two(inputs.size()));
    uint8_t r[16];
    to_le16(root.value, r);
    Device::check(zkb_ntt(Device::ctx(), r, in.data(), in.size(), out.data()));
    return attach_field(f, out, out.size());
}
// intt(root, input) ntt.rs:51-68
inline std::vector<FieldElement> intt(const FieldElement& root, const std::vector<FieldElement>& input) {
    if (input.size() < 2) return input;
    const Field* f = field_of(input, root.field);
    std::vector<u128> in = pack_values(input), out(next_power_of_two(input.size()));
    uint8_t r[16];
    to_le16(root.value, r);
    Device::check(zkb_intt(Device::ctx(), r, in.data(), in.size(), out.data()));
    return attach_field(f, out, out.size());
}

// ---- fft/ntt_arithmetics.rs ---------------------------------------------------------------------
namespace detail {
inline void assert_root_order(const FieldElement& root, u128 root_order) {      // ntt_arithmetics.rs:11-24 and siblings
    if ((root ^ root_order) != root.field->one())
        throw Panic("supplied root " + to_string(root.value) + " does not have supplied root_order " + to_string(root_order), ZKB_ERR_ROOT_ORDER);
    if ((root ^ (root_order / 2)) == root.field->one())
        throw Panic("supplied root " + to_string(root.value) + " is not a primitive of root_order " + to_string(root_order), ZKB_ERR_ROOT_ORDER);
}
}  // namespace detail

// fast_multiply ntt_arithmetics.rs:5-64 (order-shrink rule, truncation to deg l + deg r + 1, zero operand -> vec![])
inline Polynomial fast_multiply(const FieldElement& root, u128 root_order, const Polynomial& lhs, const Polynomial& rhs) {
    root.field->require_prime();
    std::vector<u128> l = pack_values(lhs.coefficients), r = pack_values(rhs.coefficients), out(std::max<size_t>(l.size() + r.size(), 1));
    uint8_t w[16];
    to_le16(root.value, w);
    size_t n_out = 0;
    Device::check(zkb_poly_mul(Device::ctx(), w, (uint64_t)root_order, l.data(), l.size(), r.data(), r.size(), out.data(), &n_out));
    return Polynomial(attach_field(root.field, out, n_out));
}
// fast_coset_evaluate ntt_arithmetics.rs:161-170 = the LDE: evaluations on offset * <generator>
inline std::vector<FieldElement> fast_coset_evaluate(const FieldElement& generator, u128 root_order, const FieldElement& offset,
                                                     const Polynomial& polynomial) {
    generator.field->require_prime();
    std::vector<u128> c = pack_values(polynomial.coefficients), out((size_t)root_order);
    uint8_t w[16], o[16];
    to_le16(generator.value, w);
    to_le16(offset.value, o);
    Device::check(zkb_coset_lde(Device::ctx(), w, (uint64_t)root_order, o, c.data(), c.size(), out.data()));
    return attach_field(generator.field, out, out.size());
}
// fast_coset_divide ntt_arithmetics.rs:239-310
inline Polynomial fast_coset_divide(const FieldElement& root, u128 root_order, const FieldElement& offset, const Polynomial& lhs,
                                    const Polynomial& rhs) {
    root.field->require_prime();
    std::vector<u128> l = pack_values(lhs.coefficients), r = pack_values(rhs.coefficients), out(std::max<size_t>(l.size(), 1));
    uint8_t w[16], o[16];
    to_le16(root.value, w);
    to_le16(offset.value, o);
    size_t n_out = 0;
    Device::check(zkb_coset_div(Device::ctx(), w, (uint64_t)root_order, o, l.data(), l.size(), r.data(), r.size(), out.data(), &n_out));
    return Polynomial(attach_field(root.field, out, n_out));
}
// The three divide-and-conquer algorithms over arbitrary domains (ntt_arithmetics.rs:66-237): the reference's
// recursion; its fast_multiply calls run on the GPU, the scalar `% * + evaluate /` glue stays scalar as there.
namespace detail {
using FEs = std::vector<FieldElement>;
inline FEs slice(const FEs& v, size_t a, size_t b) { return FEs(v.begin() + (long)a, v.begin() + (long)b); }
inline Polynomial zerofier_inner(const FieldElement& root, u128 order, const FEs& domain) {
    if (domain.empty()) return Polynomial();
    if (domain.size() == 1) return Polynomial({-domain[0], root.field->one()});
    size_t half = domain.size() / 2;
    return fast_multiply(root, order, zerofier_inner(root, order, slice(domain, 0, half)),
                         zerofier_inner(root, order, slice(domain, half, domain.size())));
}
}  // namespace detail
inline Polynomial fast_zerofier(const FieldElement& root, u128 root_order, const std::vector<FieldElement>& domain) {
    detail::assert_root_order(root, root_order);
    return detail::zerofier_inner(root, root_order, domain);
}
namespace detail {
inline FEs evaluate_inner(const FieldElement& root, u128 order, const Polynomial& p, const FEs& domain) {
    if (domain.empty()) return {};
    if (domain.size() == 1) return {p.evaluate(domain[0])};
    size_t half = domain.size() / 2;
    FEs lo = slice(domain, 0, half), hi = slice(domain, half, domain.size());
    Polynomial lz = fast_zerofier(root, order, lo), rz = fast_zerofier(root, order, hi);
    FEs left = evaluate_inner(root, order, p % lz, lo), right = evaluate_inner(root, order, p % rz, hi);
    left.insert(left.end(), right.begin(), right.end());
    return left;
}
}  // namespace detail
inline std::vector<FieldElement> fast_evaluate_domain(const FieldElement& root, u128 root_order, const Polynomial& polynomial,
                                                      const std::vector<FieldElement>& domain) {
    detail::assert_root_order(root, root_order);
    return detail::evaluate_inner(root, root_order, polynomial, domain);
}
namespace detail {
inline Polynomial interpolate_inner(const FieldElement& root, u128 order, const FEs& domain, const FEs& values) {
    if (domain.empty()) return Polynomial();
    if (domain.size() == 1) return Polynomial({values[0]});
    size_t half = domain.size() / 2;
    FEs lo = slice(domain, 0, half), hi = slice(domain, half, domain.size());
    Polynomial lz = fast_zerofier(root, order, lo), rz = fast_zerofier(root, order, hi);
    FEs lo_off = fast_evaluate_domain(root, order, rz, lo), hi_off = fast_evaluate_domain(root, order, lz, hi);
    FEs lt, rt;
    for (size_t i = 0; i < lo_off.size(); i++) lt.push_back(values[i] / lo_off[i]);
    for (size_t i = 0; i < hi_off.size(); i++) rt.push_back(values[i + half] / hi_off[i]);
    Polynomial li = interpolate_inner(root, order, lo, lt), ri = interpolate_inner(root, order, hi, rt);
    return li * rz + ri * lz;
}
}  // namespace detail
inline Polynomial fast_interpolate_domain(const FieldElement& root, u128 root_order, const std::vector<FieldElement>& domain,
                                          const std::vector<FieldElement>& values) {
    detail::assert_root_order(root, root_order);
    if (domain.size() != values.size()) throw Panic("assertion failed: domain.len() == values.len()");
    return detail::interpolate_inner(root, root_order, domain, values);
}

// ---- merkle_root.rs (T = FieldElement, the only T the crate uses on this path) -------------------
struct MerkleRoot {
    static Bytes commit(const std::vector<FieldElement>& leafs) {                       // merkle_root.rs:21-32
        if (leafs.empty()) throw Panic("length must be power of two", ZKB_ERR_NOT_POW2);  // merkle_root.rs:9 (0 is not a power of two)
        field_of(leafs);
        std::vector<u128> v = pack_values(leafs);
        Bytes root = Bytes::zeroed(64);
        Device::check(zkb_merkle_commit(Device::ctx(), v.data(), v.size(), root.buf.data()));
        return root;
    }
    // merkle_root.rs:55-66.  The reference rebuilds the tree per call; open_many opens any number of indices from one build.
    static std::vector<Bytes> open(size_t index, const std::vector<FieldElement>& leafs) { return open_many({(uint64_t)index}, leafs)[0]; }
    static std::vector<std::vector<Bytes>> open_many(const std::vector<uint64_t>& indices, const std::vector<FieldElement>& leafs) {
        if (leafs.empty()) throw Panic("length must be power of two", ZKB_ERR_NOT_POW2);  // merkle_root.rs:36
        field_of(leafs);
        std::vector<u128> v = pack_values(leafs);
        zkb_tree* t = nullptr;
        Device::check(zkb_merkle_build(Device::ctx(), v.data(), v.size(), &t));
        size_t depth = 0;
        while (((size_t)1 << depth) < v.size()) depth++;
        std::vector<uint8_t> raw(std::max<size_t>(indices.size() * depth * 64, 1));
        int rc = zkb_merkle_open(t, indices.data(), indices.size(), raw.data());
        zkb_merkle_free(t);
        Device::check(rc);
        std::vector<std::vector<Bytes>> out(indices.size());
        for (size_t s = 0; s < indices.size(); s++)
            for (size_t l = 0; l < depth; l++) out[s].emplace_back(raw.data() + (s * depth + l) * 64, 64);
        return out;
    }
    static bool verify(const Bytes& root, size_t index, const std::vector<Bytes>& path, const FieldElement& leaf) {   // merkle_root.rs:69-95
        if (root.buf.size() != 64) return false;
        std::vector<uint8_t> flat;
        for (const Bytes& b : path) {
            if (b.buf.size() != 64) return false;
            flat.insert(flat.end(), b.buf.begin(), b.buf.end());
        }
        uint8_t l[16];
        to_le16(leaf.value, l);
        return zkb_merkle_verify(root.buf.data(), index, flat.data(), path.size(), l) == 1;
    }
};

// ---- stark/proof_stream_enum.rs -------------------------------------------------------------------
struct StarkProofStreamEnum {
    enum Kind : uint8_t { Root = 0, Codeword = 1, Path = 2, Leafs = 3, Value = 4 } kind = Root;
    Bytes root;
    std::vector<FieldElement> codeword;
    std::vector<Bytes> path;
    std::array<FieldElement, 3> leafs;
    FieldElement value;
    static StarkProofStreamEnum Root_(Bytes b) { StarkProofStreamEnum o; o.kind = Root; o.root = std::move(b); return o; }
    static StarkProofStreamEnum Codeword_(std::vector<FieldElement> c) { StarkProofStreamEnum o; o.kind = Codeword; o.codeword = std::move(c); return o; }
    static StarkProofStreamEnum Path_(std::vector<Bytes> p) { StarkProofStreamEnum o; o.kind = Path; o.path = std::move(p); return o; }
    static StarkProofStreamEnum Leafs_(FieldElement a, FieldElement b, FieldElement c) { StarkProofStreamEnum o; o.kind = Leafs; o.leafs = {a, b, c}; return o; }
    static StarkProofStreamEnum Value_(FieldElement v) { StarkProofStreamEnum o; o.kind = Value; o.value = v; return o; }
    Bytes expect_root() const { need(Root, "expected to receive root"); return root; }                       // proof_stream_enum.rs:129-158
    std::vector<FieldElement> expect_codeword() const { need(Codeword, "expected to receive codeword"); return codeword; }
    std::vector<Bytes> expect_path() const { need(Path, "expected to receive path"); return path; }
    std::array<FieldElement, 3> expect_leafs() const { need(Leafs, "expected to receive leafs"); return leafs; }
    FieldElement expect_value() const { need(Value, "expected to receive value"); return value; }
    bool operator==(const StarkProofStreamEnum& o) const {
        return kind == o.kind && root == o.root && codeword == o.codeword && path == o.path && leafs == o.leafs && value == o.value;
    }

private:
    void need(Kind k, const char* msg) const { if (kind != k) throw Panic(msg); }
};

// ---- proof_stream.rs, rescue_prime/proof_stream.rs ---------------------------------------------------
// trait ProofStream<StarkProofStreamEnum>
struct ProofStream {
    virtual ~ProofStream() = default;
    virtual Bytes digest() const = 0;
    virtual Bytes fiat_shamir_prover(size_t num_bytes) const = 0;
    virtual Bytes fiat_shamir_verifier(size_t num_bytes) const = 0;
    virtual void push(const StarkProofStreamEnum& obj) = 0;
    virtual std::optional<StarkProofStreamEnum> pull() = 0;
    // The library's own stream behind this one, if any: lets FRI::prove run as ONE native call
    // (zkb_fri_prove) instead of a Fiat-Shamir callback per round.  Foreign streams return nullptr.
    virtual zkb_ps* native() { return nullptr; }
    virtual void native_appended(const Field*) {}
};

namespace detail {
inline void ps_push(zkb_ps* h, const StarkProofStreamEnum& o) {
    uint8_t a[16], b[16], c[16];
    switch (o.kind) {
        case StarkProofStreamEnum::Root: zkb_ps_push_root(h, o.root.buf.data(), o.root.buf.size()); break;
        case StarkProofStreamEnum::Codeword: {
            std::vector<u128> v = pack_values(o.codeword);
            zkb_ps_push_codeword(h, v.data(), v.size());
            break;
        }
        case StarkProofStreamEnum::Path: {
            bool digests = true;
            for (const Bytes& n : o.path) digests = digests && n.buf.size() == 64;
            std::vector<uint8_t> flat;
            if (digests) {                       // the path's real shape: 64-byte nodes
                for (const Bytes& n : o.path) flat.insert(flat.end(), n.buf.begin(), n.buf.end());
                zkb_ps_push_path(h, flat.data(), o.path.size());
            } else {                             // any node size (proof_stream_enum.rs:98-104): len u64_be || bytes, per node
                for (const Bytes& n : o.path) {
                    for (int i = 7; i >= 0; i--) flat.push_back((uint8_t)((uint64_t)n.buf.size() >> (8 * i)));
                    flat.insert(flat.end(), n.buf.begin(), n.buf.end());
                }
                zkb_ps_push_object(h, 2, flat.data(), flat.size());
            }
            break;
        }
        case StarkProofStreamEnum::Leafs:
            to_le16(o.leafs[0].value, a); to_le16(o.leafs[1].value, b); to_le16(o.leafs[2].value, c);
            zkb_ps_push_leafs(h, a, b, c);
            break;
        case StarkProofStreamEnum::Value: to_le16(o.value.value, a); zkb_ps_push_value(h, a); break;
        default: throw Panic("Unknown code");
    }
}
inline uint64_t be64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; return v; }
inline u128 be128(const uint8_t* p) { u128 v = 0; for (int i = 0; i < 16; i++) v = (v << 8) | p[i]; return v; }
}  // namespace detail

// The wire format back into objects (Stark::deser_independent_proof_stream stark.rs:213-260,
// StarkProofStreamEnum::from_bytes proof_stream_enum.rs:17-65): order_be16 || (code u8 || len u64_be || payload)*
inline std::vector<StarkProofStreamEnum> deserialize_proof(const Bytes& proof, const Field* field) {
    const std::vector<uint8_t>& b = proof.buf;
    std::vector<StarkProofStreamEnum> out;
    if (b.size() < 16) throw Panic("proof shorter than its header");
    size_t o = 16;
    while (o < b.size()) {
        if (o + 9 > b.size()) throw Panic("truncated object header");
        uint8_t code = b[o];
        uint64_t len = detail::be64(&b[o + 1]);
        o += 9;
        if (len > b.size() - o) throw Panic("truncated object payload");
        const uint8_t* p = &b[o];
        switch (code) {
            case 0: out.push_back(StarkProofStreamEnum::Root_(Bytes(p, len))); break;
            case 1: {
                if (len % 16) throw Panic("incorrect size");
                std::vector<FieldElement> c;
                for (uint64_t i = 0; i < len; i += 16) c.emplace_back(field, detail::be128(p + i));
                out.push_back(StarkProofStreamEnum::Codeword_(std::move(c)));
                break;
            }
            case 2: {
                std::vector<Bytes> path;
                uint64_t q = 0;
                while (q < len) {
                    if (q + 8 > len) throw Panic("truncated path node");
                    uint64_t l = detail::be64(p + q);
                    q += 8;
                    if (l > len - q) throw Panic("truncated path node");
                    path.emplace_back(p + q, l);
                    q += l;
                }
                out.push_back(StarkProofStreamEnum::Path_(std::move(path)));
                break;
            }
            case 3:
                if (len != 48) throw Panic("incorrect size");
                out.push_back(StarkProofStreamEnum::Leafs_(FieldElement(field, detail::be128(p)), FieldElement(field, detail::be128(p + 16)),
                                                           FieldElement(field, detail::be128(p + 32))));
                break;
            case 4:
                if (len != 16) throw Panic("incorrect size");
                out.push_back(StarkProofStreamEnum::Value_(FieldElement(field, detail::be128(p))));
                break;
            default: throw Panic("Unknown code");
        }
        o += len;
    }
    return out;
}

// IndependentProofStream<StarkProofStreamEnum> (proof_stream.rs:14-83).  The wire bytes live in the
// library's stream (zkb_ps, hosthash.cpp: incremental serialisation) and are mirrored as objects for pull().
class IndependentProofStream : public ProofStream {
public:
    std::vector<StarkProofStreamEnum> objects;
    size_t read_index = 0;
    IndependentProofStream() { create(nullptr, 0, 0); }
    explicit IndependentProofStream(const std::vector<StarkProofStreamEnum>& objs) : IndependentProofStream() {      // ::from
        for (auto& o : objs) push(o);
    }
    ~IndependentProofStream() override { if (h_) zkb_ps_free(h_); }
    IndependentProofStream(const IndependentProofStream&) = delete;
    IndependentProofStream& operator=(const IndependentProofStream&) = delete;
    Bytes digest() const override {
        size_t n = zkb_ps_digest(h_, nullptr, 0);
        Bytes out = Bytes::zeroed(n);
        zkb_ps_digest(h_, out.buf.data(), n);
        return out;
    }
    Bytes fiat_shamir_prover(size_t num_bytes) const override {
        Bytes out = Bytes::zeroed(num_bytes);
        zkb_ps_fiat_shamir(h_, num_bytes, out.buf.data());
        return out;
    }
    Bytes fiat_shamir_verifier(size_t num_bytes) const override {               // only the objects already pulled
        zkb_ps* t = nullptr;
        if (zkb_ps_create(doc_.empty() ? nullptr : doc_.data(), doc_.size(), signature_, &t) != ZKB_OK) throw Panic("zkb_ps_create failed");
        for (size_t i = 0; i < read_index; i++) detail::ps_push(t, objects[i]);
        Bytes out = Bytes::zeroed(num_bytes);
        zkb_ps_fiat_shamir(t, num_bytes, out.buf.data());
        zkb_ps_free(t);
        return out;
    }
    void push(const StarkProofStreamEnum& obj) override {
        detail::ps_push(h_, obj);
        objects.push_back(obj);
    }
    std::optional<StarkProofStreamEnum> pull() override {
        if (read_index >= objects.size()) throw Panic("Cannot pull, queue is empty");
        return objects[read_index++];
    }
    zkb_ps* native() override { return h_; }
    void native_appended(const Field* field) override { objects = deserialize_proof(digest(), field); }
    bool operator==(const IndependentProofStream& o) const { return objects == o.objects; }

protected:
    IndependentProofStream(const uint8_t* doc, size_t len) { doc_.assign(doc, doc + len); signature_ = 1; create(doc, len, 1); }

private:
    zkb_ps* h_ = nullptr;
    std::vector<uint8_t> doc_;
    int signature_ = 0;
    void create(const uint8_t* doc, size_t len, int sig) {
        if (zkb_ps_create(doc, len, sig, &h_) != ZKB_OK || !h_) throw Panic("zkb_ps_create failed");
    }
};

// SignatureProofStream (rescue_prime/proof_stream.rs:9-62): the Fiat-Shamir input is prefixed with
// u64_be(64) || BLAKE2b-512(document); the stored proof is not.
class SignatureProofStream : public IndependentProofStream {
public:
    explicit SignatureProofStream(const Bytes& document) : IndependentProofStream(document.buf.data(), document.buf.size()) {}
};

// ---- fri.rs -----------------------------------------------------------------------------------------
struct Result {                                      // Result<(), String>
    std::optional<std::string> err;
    bool is_ok() const { return !err.has_value(); }
    static Result Ok() { return Result(); }
    static Result Err(std::string m) { Result r; r.err = std::move(m); return r; }
    bool operator==(const Result& o) const { return err == o.err; }
    bool operator!=(const Result& o) const { return !(*this == o); }
};

class FRI {
public:
    FieldElement omega, offset;
    const Field* field;
    size_t domain_length, expansion_factor, num_colinearity_tests;
    FRI(const FieldElement& offset_, const FieldElement& omega_, size_t domain_length_, size_t expansion_factor_, size_t num_colinearity_tests_)
        : omega(omega_), offset(offset_), field(omega_.field), domain_length(domain_length_), expansion_factor(expansion_factor_),
          num_colinearity_tests(num_colinearity_tests_) {}                                               // fri.rs:23-38

    size_t num_rounds() const { zkb_fri_params p = params(); return (size_t)zkb_fri_num_rounds(&p); }    // fri.rs:40-50
    std::vector<FieldElement> evaluate_domain() const {                                                  // fri.rs:52-58 (only its length is used by Stark)
        std::vector<FieldElement> out;
        out.reserve(domain_length);
        FieldElement x = offset;
        for (size_t i = 0; i < domain_length; i++) { out.push_back(x); x = x * omega; }
        return out;
    }
    std::vector<size_t> sample_indices(const Bytes& seed, size_t size, size_t reduced_size, size_t number) const {   // fri.rs:85-113
        if (number > reduced_size) throw Panic("Cannot sample more indices than available in the last codeword");
        std::vector<uint64_t> out(number);
        if (zkb_fri_sample_indices(seed.buf.data(), seed.buf.size(), size, reduced_size, number, out.data()) != ZKB_OK)
            throw Panic("sample_indices failed");
        return std::vector<size_t>(out.begin(), out.end());
    }

    // FRI::prove fri.rs:210-248 -> top-level indices.  Commit (Merkle trees, fused fold + leaf hashing) and
    // the query openings run on the GPU; the layers never leave HBM.
    std::vector<size_t> prove(const std::vector<FieldElement>& codeword, ProofStream& proof_stream) const {
        if (codeword.size() != domain_length) throw Panic("Length of the domain doesnt match the length of initial codeword", ZKB_ERR_LENGTH);
        field->require_prime();
        zkb_fri_params p = params();
        std::vector<u128> cw = pack_values(codeword);
        if (zkb_ps* h = proof_stream.native()) {                       // the library's own stream: one call
            std::vector<uint64_t> top(num_colinearity_tests);
            Device::check(zkb_fri_prove(Device::ctx(), &p, cw.data(), cw.size(), h, top.data()));
            proof_stream.native_appended(field);
            return std::vector<size_t>(top.begin(), top.end());
        }
        // any other ProofStream implementation: Fiat-Shamir through a callback per round (fri.rs:136-146)
        struct Hop { ProofStream* ps; const Field* f; std::string err; } hop{&proof_stream, field, {}};
        zkb_fs_callback cb = [](void* user, uint32_t, const uint8_t root[64], int want_alpha, uint8_t alpha_out[16]) -> int {
            Hop* hp = static_cast<Hop*>(user);
            try {
                hp->ps->push(StarkProofStreamEnum::Root_(Bytes(root, 64)));
                if (want_alpha) to_le16(hp->f->sample(hp->ps->fiat_shamir_prover(PROOF_BYTES)).value, alpha_out);
                return 0;
            } catch (const std::exception& e) { hp->err = e.what(); return 1; }
        };
        zkb_fri_layers* layers = nullptr;
        int rc = zkb_fri_commit(Device::ctx(), &p, cw.data(), cw.size(), cb, &hop, &layers);
        if (!hop.err.empty()) throw Panic(hop.err, ZKB_ERR_CALLBACK);
        Device::check(rc);
        std::unique_ptr<zkb_fri_layers, void (*)(zkb_fri_layers*)> guard(layers, zkb_fri_layers_free);
        const uint64_t R = zkb_fri_layer_count(layers);
        if (R < 2) throw Panic("FRI::prove needs at least two rounds (fri.rs:225 unwraps codewords.get(1))", ZKB_ERR_ROUNDS);
        {
            std::vector<u128> last(zkb_fri_layer_len(layers, R - 1));
            Device::check(zkb_fri_layer_codeword(layers, R - 1, last.data()));
            proof_stream.push(StarkProofStreamEnum::Codeword_(attach_field(field, last, last.size())));     // fri.rs:166
        }
        std::vector<size_t> top = sample_indices(proof_stream.fiat_shamir_prover(PROOF_BYTES), zkb_fri_layer_len(layers, 1),
                                                 zkb_fri_layer_len(layers, R - 1), num_colinearity_tests);
        std::vector<uint64_t> idx(top.begin(), top.end());
        const size_t ncc = num_colinearity_tests;
        for (uint64_t r = 0; r + 1 < R; r++) {                                                               // fri.rs:231-245, 174-208
            const uint64_t len = zkb_fri_layer_len(layers, r);
            for (auto& i : idx) i %= len / 2;
            size_t d_cur = 0, d_nxt;
            while (((uint64_t)1 << d_cur) < len) d_cur++;
            d_nxt = d_cur - 1;
            std::vector<uint8_t> leafs(ncc * 48), paths(ncc * (2 * d_cur + d_nxt) * 64);
            Device::check(zkb_fri_query(layers, r, idx.data(), ncc, leafs.data(), paths.data()));
            for (size_t s = 0; s < ncc; s++)
                proof_stream.push(StarkProofStreamEnum::Leafs_(FieldElement(field, from_le16(&leafs[48 * s])), FieldElement(field, from_le16(&leafs[48 * s + 16])),
                                                               FieldElement(field, from_le16(&leafs[48 * s + 32]))));
            size_t o = 0;
            for (size_t s = 0; s < ncc; s++)
                for (size_t d : {d_cur, d_cur, d_nxt}) {
                    std::vector<Bytes> path;
                    for (size_t l = 0; l < d; l++, o += 64) path.emplace_back(&paths[o], 64);
                    proof_stream.push(StarkProofStreamEnum::Path_(std::move(path)));
                }
        }
        return top;
    }

    // FRI::verify fri.rs:250-416: scalar host code as in the reference (Merkle commit of the last codeword,
    // its iNTT / NTT and the scale run on the GPU).
    Result verify(ProofStream& proof_stream, std::vector<std::pair<size_t, FieldElement>>& polynomial_values) const {
        FieldElement om = omega, off = offset;
        const size_t R = num_rounds();
        std::vector<Bytes> roots;
        std::vector<FieldElement> alphas;
        for (size_t r = 0; r < R; r++) {
            roots.push_back(proof_stream.pull()->expect_root());
            alphas.push_back(field->sample(proof_stream.fiat_shamir_verifier(PROOF_BYTES)));
        }
        std::vector<FieldElement> last_codeword = proof_stream.pull()->expect_codeword();
        if (MerkleRoot::commit(last_codeword) != roots.back()) return Result::Err("last codeword is not well formed");
        const size_t degree = last_codeword.size() / expansion_factor - 1;
        FieldElement last_omega = om, last_offset = off;
        for (size_t i = 0; i + 1 < R; i++) { last_omega = last_omega ^ 2; last_offset = last_offset ^ 2; }
        if (last_omega.inverse() != (last_omega ^ (u128)(last_codeword.size() - 1))) return Result::Err("omega does not have the right order");
        Polynomial poly = Polynomial(intt(last_omega, last_codeword)).scale(last_offset.inverse());
        auto pd = poly.degree();
        if (!pd) return Result::Err("Received none instead of polynomial degree");
        if (*pd > degree)
            return Result::Err("last codeword does not correspond to polynomial of low enough degree (it is " + std::to_string(*pd) +
                               " but should be <= " + std::to_string(degree) + ")");
        if (ntt(last_omega, poly.scale(last_offset).coefficients) != last_codeword) return Result::Err("re-evaluated codeword does not match original");
        std::vector<size_t> top = sample_indices(proof_stream.fiat_shamir_verifier(PROOF_BYTES), domain_length >> 1, domain_length >> (R - 1),
                                                 num_colinearity_tests);
        for (size_t r = 0; r + 1 < R; r++) {
            const size_t half = domain_length >> (r + 1);
            std::vector<size_t> ia, ib;
            for (size_t i : top) { ia.push_back(i % half); ib.push_back(i % half + half); }
            const std::vector<size_t>& ic = ia;
            std::vector<FieldElement> aa, bb, cc;
            for (size_t s = 0; s < num_colinearity_tests; s++) {
                auto l = proof_stream.pull()->expect_leafs();
                aa.push_back(l[0]); bb.push_back(l[1]); cc.push_back(l[2]);
                if (r == 0) { polynomial_values.emplace_back(ia[s], l[0]); polynomial_values.emplace_back(ib[s], l[1]); }
                FieldElement ax = off * (om ^ (u128)ia[s]), bx = off * (om ^ (u128)ib[s]), cx = alphas[r];
                if (!Polynomial::test_colinearity({{ax, l[0]}, {bx, l[1]}, {cx, l[2]}})) return Result::Err("colinearity check failure");
            }
            for (size_t i = 0; i < num_colinearity_tests; i++) {
                if (!MerkleRoot::verify(roots[r], ia[i], proof_stream.pull()->expect_path(), aa[i])) return Result::Err("Merkle auth path verification failed for aa");
                if (!MerkleRoot::verify(roots[r], ib[i], proof_stream.pull()->expect_path(), bb[i])) return Result::Err("Merkle auth path verification failed for bb");
                if (!MerkleRoot::verify(roots[r + 1], ic[i], proof_stream.pull()->expect_path(), cc[i])) return Result::Err("Merkle auth path verification failed for cc");
            }
            om = om ^ 2;
            off = off ^ 2;
        }
        return Result::Ok();
    }

private:
    zkb_fri_params params() const {
        zkb_fri_params p;
        to_le16(offset.value, p.offset);
        to_le16(omega.value, p.omega);
        p.domain_length = domain_length;
        p.expansion_factor = expansion_factor;
        p.num_colinearity_tests = num_colinearity_tests;
        return p;
    }
};

}  // namespace zk_impl
#endif  // ZK_IMPL_HPP
