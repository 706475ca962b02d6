// The reference crate's own #[test] functions for the hot path, restated in C++ against
// include/zk_impl.hpp (the C++ host side of the drop-in) - same inputs, same expected constants,
// same assertions, so this file reads like the crate's test modules:
//   src/field/field.rs:176-256, src/field/field_element.rs:150-299, src/crypto/blake2b512.rs:20-31,
//   src/proof_stream.rs:87-146 (SHAKE challenges), src/stark/stark.rs:785-808 (wire format),
//   src/fft/ntt.rs:78-130, src/field/polynomial.rs:631-652 (scale), src/fft/ntt_arithmetics.rs:355-517,
//   src/merkle_root.rs:107-244, src/fri.rs:426-531.
// `reference_tests host` runs the groups that need no GPU; `reference_tests` (or `all`) runs everything.
// The reference draws random polynomials from thread_rng; here they come from a fixed xorshift so a
// failure reproduces.  Sub-assertions over Field::new(100) / Field::new(8) are outside the GPU path's
// precondition (FIELD_PRIME only) and only their add/sub/neg parts (pure host) are kept.
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>

#include "zk_impl.hpp"

using namespace zk_impl;

static int g_fail = 0, g_run = 0;
#define ASSERT(cond, ...)                                                              \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            std::printf("    ASSERT FAILED %s:%d: %s : ", __FILE__, __LINE__, #cond);  \
            std::printf(__VA_ARGS__);                                                  \
            std::printf("\n");                                                         \
            throw std::runtime_error("assertion failed");                              \
        }                                                                              \
    } while (0)
#define ASSERT_EQ(a, b) ASSERT((a) == (b), "left != right")
#define ASSERT_NE(a, b) ASSERT((a) != (b), "left == right")

static void run(const char* name, const std::function<void()>& f) {
    g_run++;
    try {
        f();
        std::printf("test %s ... ok\n", name);
    } catch (const std::exception& e) {
        g_fail++;
        std::printf("test %s ... FAILED (%s)\n", name, e.what());
    }
    std::fflush(stdout);
}

static FieldElement fe(const Field& f, const char* dec) { return FieldElement(&f, parse_u128(dec)); }
static FieldElement fe(const Field& f, unsigned long long v) { return FieldElement(&f, (u128)v); }
static FieldElement fe(const Field& f, int v) { return FieldElement(&f, (u128)v); }
static FieldElement fe(const Field& f, size_t v) { return FieldElement(&f, (u128)v); }
static std::vector<FieldElement> fes(const Field& f, std::initializer_list<const char*> decs) {
    std::vector<FieldElement> out;
    for (const char* d : decs) out.push_back(fe(f, d));
    return out;
}
static std::vector<FieldElement> fes(const Field& f, std::initializer_list<unsigned long long> vs) {
    std::vector<FieldElement> out;
    for (auto v : vs) out.push_back(fe(f, v));
    return out;
}

// deterministic stand-in for thread_rng (ntt_arithmetics.rs:321-353: rand_domain / rand_poly)
static uint64_t g_rng = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd64() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }
static std::vector<FieldElement> rand_domain(const Field& f, size_t n) {
    std::vector<FieldElement> out;
    for (size_t i = 0; i < n; i++) out.emplace_back(&f, ((((u128)rnd64()) << 64) | rnd64()) % FIELD_PRIME);
    return out;
}
static Polynomial rand_poly(const Field& f, size_t max_degree) {
    size_t degree = 0;
    while (degree == 0) degree = (size_t)(rnd64() & 0xFF) % max_degree;
    return Polynomial(rand_domain(f, degree));
}

static int objs_kind(const std::vector<StarkProofStreamEnum>& o, size_t i) { return (int)o.at(i).kind; }

// ------------------------------------------------------------------------------------------- host
static void host_tests() {
    run("field::field::tests::mul", [] {                                  // field.rs:176-183 (through FieldElement)
        Field field(FIELD_PRIME);
        ASSERT_EQ((fe(field, 2) * fe(field, 3)).value, (u128)6);
        ASSERT_EQ((FieldElement(&field, FIELD_PRIME - 1) * fe(field, 3)).value, FIELD_PRIME - 3);
    });
    run("field::field::tests::primitive_nth_root", [] {                   // field.rs:185-217
        Field field(FIELD_PRIME);
        const u128 n = 256;
        const int n_log = 8;
        FieldElement z = field.primitive_nth_root(n);
        ASSERT_EQ(field.primitive_nth_root(256), fe(field, "178902808384765167578311106676137348214"));
        ASSERT_EQ(field.primitive_nth_root(2), fe(field, "270497897142230380135924736767050121216"));
        FieldElement powered = z;
        for (u128 i = 0; i < n - 1; i++) powered = powered * z;
        ASSERT(powered.value == 1, "omega is not 256th root of unity");
        powered = z;
        for (u128 i = 0; i < n - 2; i++) powered = powered * z;
        ASSERT(powered.value != 1, "omega is not primitive");
        ASSERT_EQ(z ^ ((u128)1 << n_log), field.one());
        ASSERT_NE(z ^ ((u128)1 << (n_log - 1)), field.one());
    });
    run("field::field::tests::sample", [] {                               // field.rs:219-241
        Field field(FIELD_PRIME);
        ASSERT_EQ(field.sample(Bytes("6c9c4992")), fe(field, 1822181778ULL));
        ASSERT_EQ(field.sample(Bytes("ac4cd3be")), fe(field, 2890716094ULL));
    });
    run("field::field::tests::neg", [] {                                  // field.rs:243-256
        Field field(FIELD_PRIME);
        ASSERT_EQ((-fe(field, 256)).value, parse_u128("270497897142230380135924736767050120961"));
        Field f100(100);
        ASSERT_EQ((fe(f100, 20) + (-fe(f100, 20))).value, (u128)0);
        ASSERT_EQ((fe(f100, 20) + (-fe(f100, 19))).value, (u128)1);
    });
    run("field::field_element::tests::mul", [] {                          // field_element.rs:150-165
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, "49789714223038013592473676705012096123") * fe(field, "6534789852937546098347957826345234"),
                  fe(field, "105250150227149389100670877502232671566"));
        ASSERT_EQ(fe(field, 8) * fe(field, 12), fe(field, 96));
        ASSERT_EQ(fe(field, 3) * fe(field, "270497897142230380135924736767050121215"), fe(field, "270497897142230380135924736767050121211"));
    });
    run("field::field_element::tests::div", [] {                          // field_element.rs:167-199
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, "74658620945386735627456854792784352353") / fe(field, "85408008396924667383611388730472331217"),
                  fe(field, "120557879365253444230411244907275635216"));
        ASSERT_EQ(fe(field, 12) / fe(field, 4), fe(field, 3));
        ASSERT_EQ(fe(field, "270497897142230380135924736767050121215") / fe(field, 5), fe(field, "54099579428446076027184947353410024243"));
        ASSERT_EQ(fe(field, 5012096123ULL) / fe(field, "6534789852937546098347957826345234"), fe(field, "109071144973379706934869779239844248849"));
        bool panicked = false;
        try { (void)(fe(field, 1) / fe(field, 0)); } catch (const Panic& p) { panicked = std::string(p.what()) == "divide by zero"; }
        ASSERT(panicked, "division by zero must panic (field_element.rs:85)");
    });
    run("field::field_element::tests::inverse", [] {                      // field_element.rs:201-221
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, 256).inverse(), fe(field, "269441264731518542713518780764053831681"));
        FieldElement el = fe(field, 8);
        ASSERT_EQ(el * el.inverse(), field.one());
        el = fe(field, "270497897142230380135924736767050121215");
        ASSERT_EQ(el * el.inverse(), field.one());
    });
    run("field::field_element::tests::add", [] {                          // field_element.rs:223-244
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, "270497897142230380135924736767050120961") + fe(field, 300), fe(field, 44));
        Field f100(100);
        ASSERT_EQ(fe(f100, 20) + fe(f100, 20), fe(f100, 40));
        ASSERT_EQ(fe(f100, 20) + (-fe(f100, 19)), f100.one());
        ASSERT_EQ(fe(f100, 80) + fe(f100, 21), f100.one());
    });
    run("field::field_element::tests::sub", [] {                          // field_element.rs:246-267
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, 44) - fe(field, 200), fe(field, "270497897142230380135924736767050121061"));
        Field f100(100);
        ASSERT_EQ(fe(f100, 20) - fe(f100, 20), f100.zero());
        ASSERT_EQ(fe(f100, 20) - fe(f100, 19), f100.one());
        ASSERT_EQ(fe(f100, 20) - fe(f100, 21), -f100.one());
    });
    run("field::field_element::tests::neg", [] {                          // field_element.rs:269-286
        Field field(FIELD_PRIME);
        ASSERT_EQ(-fe(field, 6534789852937546098ULL), fe(field, "270497897142230380129389946914112575119"));
        Field f100(100);
        ASSERT_EQ(-fe(f100, 1), fe(f100, 99));
        ASSERT_EQ(-fe(f100, 20), fe(f100, 80));
    });
    run("field::field_element::tests::pow", [] {                          // field_element.rs:288-299
        Field field(FIELD_PRIME);
        ASSERT_EQ(fe(field, 6534789852937546098ULL) ^ (u128)501209126122ULL, fe(field, "256557788041265930815463337858691703671"));
        ASSERT_EQ(fe(field, 15) ^ 4, fe(field, 50625));
        ASSERT_EQ(fe(field, "270497897142230380135") ^ 8, fe(field, "79016866124691016201920330826259043252"));
    });
    run("crypto::blake2b512::tests::test", [] {                           // blake2b512.rs:20-31
        ASSERT_EQ(blake2b512(Bytes(std::vector<uint8_t>{0})).to_hex(),
                  std::string("2fa3f686df876995167e7c2e5d74c4c7b6e48f8068fe0e44208344d480f7904c36963e44115fe3eb2a3ac8694c28bcb4f5a0f3276f2e79487d8219057a506e4b"));
        ASSERT_EQ(blake2b512(Bytes(std::vector<uint8_t>{0, 0})).to_hex(),
                  std::string("5ba7f7e4ade7e5803c59d184326420823f7f860effcfba0bb896d568f59b8d85181cfff25929d40b18e01069c2ef5c31754f1d821a1f3f80f896f4dde374a2f1"));
    });
    run("proof_stream::tests::order (SHAKE256 challenges)", [] {          // proof_stream.rs:87-146: digest = Debug text of the objects
        auto text = [](const char* s) { return Bytes(reinterpret_cast<const uint8_t*>(s), std::strlen(s)); };
        ASSERT_EQ(shake256(text("[]"), 64).to_hex(),
                  std::string("ec784925b52067bce01fd820f554a34a3f8522b337f82e00ea03d3fa2b207ef9c2c1b9ed900cf2bbfcd19a232a94c6121e041615305c4155d46d52f58a8cff1c"));
        Bytes t = text("[Str(\"Hello, World!\"), Vec([0, 1, 5, 234]), Map({\"something\": 123})]");
        ASSERT_EQ(shake256(t, 4).to_hex(), std::string("78b0db5c"));
        ASSERT_EQ(shake256(t, 64).to_hex(),
                  std::string("78b0db5cfd13c78498fd0951a9fd609f2521fd02d850cc561eced844bb0c338588358abcc0d98d76c6779cb388514f4bc19e2c0125b143abee166cb98c38a831"));
    });
    run("stark::stark::tests::deserialize_proof_stream", [] {             // stark.rs:785-808
        Field field(FIELD_PRIME);
        std::vector<StarkProofStreamEnum> objs = {
            StarkProofStreamEnum::Root_(Bytes(std::vector<uint8_t>{0x49, 0x6e, 0x20, 0x74})),
            StarkProofStreamEnum::Codeword_({fe(field, 20), fe(field, 100)}),
            StarkProofStreamEnum::Path_({Bytes(std::vector<uint8_t>{0x49, 0x6e, 0x20, 0x74}), Bytes(std::vector<uint8_t>{0x1, 0x6b, 0xfe, 0x25})}),
            StarkProofStreamEnum::Leafs_(fe(field, 1), fe(field, 5), fe(field, 10)),
            StarkProofStreamEnum::Value_(fe(field, 2)),
        };
        IndependentProofStream stream(objs);
        Bytes serialized = stream.digest();
        // header (p, big-endian: a field-carrying object is present) + 5 x (code, len) + payloads, SURVEY.md A.4
        ASSERT_EQ(serialized.buf.size(), (size_t)(16 + 5 * 9 + 4 + 32 + 24 + 48 + 16));
        ASSERT_EQ(Bytes(serialized.buf.data(), 16).to_hex(), std::string("cb800000000000000000000000000001"));
        IndependentProofStream deserialized(deserialize_proof(serialized, &field));
        ASSERT(stream == deserialized, "round trip changed the objects");
        ASSERT_EQ(deserialized.digest(), serialized);
        // verifier-side challenge covers only what was pulled (proof_stream.rs:43-48); zero header before any field object
        ASSERT_EQ(stream.fiat_shamir_verifier(32), shake256(Bytes(std::vector<uint8_t>(16, 0)), 32));
        stream.pull();
        ASSERT_NE(stream.fiat_shamir_verifier(32), stream.fiat_shamir_prover(32));
        for (int i = 0; i < 4; i++) stream.pull();
        ASSERT_EQ(stream.fiat_shamir_verifier(32), stream.fiat_shamir_prover(32));
        bool panicked = false;
        try { stream.pull(); } catch (const Panic&) { panicked = true; }
        ASSERT(panicked, "pull on an exhausted stream must panic (proof_stream.rs:55)");
        // SignatureProofStream: same stored bytes, document-prefixed challenge (rescue_prime/proof_stream.rs:24-52)
        SignatureProofStream sig(Bytes(std::vector<uint8_t>{'d', 'o', 'c'}));
        for (auto& o : objs) sig.push(o);
        ASSERT_EQ(sig.digest(), serialized);
        Bytes prefix = Bytes(std::vector<uint8_t>{0, 0, 0, 0, 0, 0, 0, 64}) + blake2b512(Bytes(std::vector<uint8_t>{'d', 'o', 'c'}));
        ASSERT_EQ(sig.fiat_shamir_prover(32), shake256(prefix + serialized, 32));
        ASSERT_NE(sig.fiat_shamir_prover(32), stream.fiat_shamir_prover(32));
    });
    run("merkle_root::tests::verify", [] {                                // merkle_root.rs:204-244
        Field field(FIELD_PRIME);
        const char* root = "b36f5edab7ea2100fc298d9811bf1a745745282e80243e3a919e71ef6c30f690606b445557ad7843d3251c8e92b83b584d94b738334ffa7d88babd6e47471ac5";
        std::vector<Bytes> path = {
            "1f069c52b4f26c7714dbd9babacbff542d1333190e3246dec47ee9f30bb649046406f3e0ae8f4cafd52bc1a1305061b451a8746ad3ad240c2524a82a3fcd28c0",
            "9b70e42c4b3aea3efddaeda6c1883b38c8969e40ca17566d612156c0457961e7c30d811e2adefd941da7b5329d24ecf015dcffb3e39e379dc988564d588a2341"};
        ASSERT(MerkleRoot::verify(root, 1, path, fe(field, 456)), "Root has to be valid");
        ASSERT(!MerkleRoot::verify(root, 1, path, fe(field, 5462)), "Root has to be invalid because element is invalid");
        ASSERT(!MerkleRoot::verify(root, 0, path, fe(field, 456)), "Root has to be invalid because index is invalid");
    });
    run("fri::tests::sample_indices", [] {                                // fri.rs:426-448
        const size_t n = 256;
        Field field(FIELD_PRIME);
        FRI fri(field.generator(), field.primitive_nth_root(n), n, 4, 17);
        std::vector<size_t> sample = fri.sample_indices("d4b6e8af1114859c1c24b6496a3aef2f55a21105bc103af7e12dc3b2c101fe66", 128, 128, 17);
        ASSERT_EQ(sample, (std::vector<size_t>{40, 121, 5, 113, 97, 68, 126, 88, 26, 82, 81, 91, 93, 125, 10, 57, 48}));
        ASSERT_EQ(fri.num_rounds(), (size_t)2);                           // 256 -> 128 -> 64 (<= 4 * 17)
    });
    run("field::polynomial (host helpers: degree, evaluate, + - * %)", [] {   // polynomial.rs:46-100, 252-326
        Field field(FIELD_PRIME);
        Polynomial a(fes(field, {1ULL, 2ULL, 3ULL})), b(fes(field, {5ULL, 7ULL}));
        ASSERT_EQ((a * b), Polynomial(fes(field, {5ULL, 17ULL, 29ULL, 21ULL})));
        ASSERT_EQ(((a * b) % a).is_zero(), true);
        auto qr = Polynomial::divide_with_rem(a * b + Polynomial(fes(field, {4ULL})), a);
        ASSERT_EQ(qr.first, b);
        ASSERT_EQ(qr.second.coefficients[0], fe(field, 4));
        ASSERT_EQ(a.evaluate(fe(field, 10)), fe(field, 321));
        ASSERT_EQ(Polynomial(fes(field, {0ULL, 0ULL})).degree().has_value(), false);
        ASSERT_EQ(*Polynomial(fes(field, {0ULL, 4ULL, 0ULL})).degree(), (size_t)1);
        ASSERT(Polynomial::test_colinearity({{fe(field, 1), fe(field, 3)}, {fe(field, 2), fe(field, 5)}, {fe(field, 10), fe(field, 21)}}), "colinear points");
        ASSERT(!Polynomial::test_colinearity({{fe(field, 1), fe(field, 3)}, {fe(field, 2), fe(field, 5)}, {fe(field, 10), fe(field, 22)}}), "not colinear");
    });
}

// -------------------------------------------------------------------------------------------- GPU
static void gpu_tests() {
    run("fft::ntt::tests::test_ntt", [] {                                 // ntt.rs:78-105
        Field field(FIELD_PRIME);
        const u128 n = 1 << 4;
        FieldElement primitive_root = field.primitive_nth_root(n);
        auto input = fes(field, {"10350860596407318609598574026175964133", "60692809610834653822383343680910625982", "223446197944610152228521360138425742723", "123599176902523769876954930401435714041", "233214499950980668770362073427851594143", "197530481770421435151547222505733630031", "6028204552208455457232478170590637777", "129106051215868132791440857107220454376", "46875137253396986423834480299002499296", "40573479539486208028801437611599580111", "177627388180112816822358878396956962568", "63754231379381382860231899477157171256", "213977912421556511151382836938765186268", "247295448209556494808801789962732329479", "198078312580458497833840274756537503682", "140348661180454074099943144751461445367"});
        auto values = fes(field, {"219013573292644897785762424283206192714", "28020178707455534238013018981848447223", "125720672179066355667363683873634014638", "9544075888957995047526079628773702483", "236214009288214032104373542256167121711", "203576991437594049347129945434757211067", "161303837601531457486430204397030363075", "8066037348193233635957451882263404827", "106698173671205255857026330656055139947", "205516443913240407551582667880265743260", "132452175458644240344865798387130681692", "14403130148933356826258737037147692544", "103258398926149853393925877736501914903", "241567637358481607032146874821122458208", "184640833807669488035490783312852642403", "79102880510147994351196921219409543952"});
        ASSERT_EQ(input.size(), (size_t)n);
        ASSERT_EQ(ntt(primitive_root, input), values);
        std::vector<FieldElement> domain;
        for (size_t i = 0; i < values.size(); i++) domain.push_back(primitive_root ^ (u128)i);
        ASSERT_EQ(Polynomial(input).evaluate_domain(domain), values);
    });
    run("fft::ntt::tests::test_intt", [] {                                // ntt.rs:108-130
        Field field(FIELD_PRIME);
        FieldElement primitive_root = field.primitive_nth_root(1 << 4);
        auto values = fes(field, {159ULL, 179ULL, 197ULL, 143ULL, 198ULL, 82ULL, 100ULL, 153ULL, 45ULL, 158ULL, 154ULL, 238ULL, 46ULL, 121ULL, 148ULL, 200ULL});
        auto coeffs = fes(field, {"2321", "46679697743149797158402415879589215379", "85767599764045409871854383990500128680", "170048455543476672374689900824216177289", "56517926799859326797837626323965682333", "150718635918560071455504820257610329093", "149093701728889244918633279335367822666", "266977550113122771518657412035427200127", "270497897142230380135924736767050120990", "63434915687244166391766073758524869310", "261359683971832165794823307314869483630", "172866549451408128829178953127270691728", "213979970342371053338087110443084438582", "83513730222590766426030683639871516493", "44774808819693939686538502893362807298", "127752053889369146389468687545690486361"});
        ASSERT_EQ(ntt(primitive_root, values), coeffs);
        ASSERT_EQ(intt(primitive_root, coeffs), values);
        bool panicked = false;
        try { ntt(primitive_root, {}); } catch (const Panic& p) { panicked = p.code == ZKB_ERR_EMPTY; }
        ASSERT(panicked, "ntt of an empty vector must panic (ntt.rs:11)");
        ASSERT_EQ(intt(primitive_root, {values[0]}), (std::vector<FieldElement>{values[0]}));     // len < 2: returned as is (ntt.rs:55-57)
    });
    run("field::polynomial::tests::scale", [] {                           // polynomial.rs:631-652
        Field field(FIELD_PRIME);
        Polynomial poly(fes(field, {10ULL, 345ULL, 0ULL, 65ULL, 74ULL, 5ULL}));
        Polynomial want(fes(field, {10ULL, 1380ULL, 0ULL, 4160ULL, 18944ULL, 5120ULL}));
        ASSERT_EQ(poly.scale(fe(field, 4)), want);
    });
    run("fft::ntt_arithmetics::tests::multiply", [] {                     // ntt_arithmetics.rs:355-375
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        for (int trial = 0; trial < 20; trial++) {
            Polynomial lhs = rand_poly(field, n / 2), rhs = rand_poly(field, n / 2);
            ASSERT(fast_multiply(primitive_root, n, lhs, rhs) == lhs * rhs, "#%d trial failed", trial);
        }
        ASSERT_EQ(fast_multiply(primitive_root, n, Polynomial(), rand_poly(field, 8)), Polynomial());
        bool panicked = false;
        try { fast_multiply(primitive_root, n / 2, rand_poly(field, 8), rand_poly(field, 8)); } catch (const Panic& p) { panicked = p.code == ZKB_ERR_ROOT_ORDER; }
        ASSERT(panicked, "a root of the wrong order must panic (ntt_arithmetics.rs:11-24)");
    });
    run("fft::ntt_arithmetics::tests::zerofier", [] {                     // ntt_arithmetics.rs:377-403
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        for (int trial = 0; trial < 20; trial++) {
            Polynomial poly = rand_poly(field, n);
            Polynomial zerofier = fast_zerofier(primitive_root, n, poly.coefficients);
            for (const FieldElement& c : poly.coefficients) ASSERT(zerofier.evaluate(c) == field.zero(), "#%d trial failed", trial);
        }
    });
    run("fft::ntt_arithmetics::tests::evaluate_domain", [] {              // ntt_arithmetics.rs:405-432
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        for (int trial = 0; trial < 20; trial++) {
            Polynomial poly = rand_poly(field, n);
            auto domain = rand_domain(field, n);
            ASSERT(poly.evaluate_domain(domain) == fast_evaluate_domain(primitive_root, n, poly, domain), "#%d trial failed", trial);
        }
    });
    run("fft::ntt_arithmetics::tests::interpolate", [] {                  // ntt_arithmetics.rs:434-470 (the reference runs 20 trials of 64 points)
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        for (int trial = 0; trial < 4; trial++) {
            auto domain = rand_domain(field, n), values = rand_domain(field, n);
            Polynomial poly = fast_interpolate_domain(primitive_root, n, domain, values);
            ASSERT(values == fast_evaluate_domain(primitive_root, n, poly, domain), "#%d trial failed", trial);
        }
    });
    run("fft::ntt_arithmetics::tests::coset_evaluate", [] {               // ntt_arithmetics.rs:472-492
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        FieldElement offset = fe(field, 5);
        std::vector<FieldElement> domain;
        for (u128 i = 0; i < n; i++) domain.push_back((primitive_root ^ i) * offset);
        Polynomial poly = rand_poly(field, n);
        ASSERT_EQ(fast_evaluate_domain(primitive_root, n, poly, domain), fast_coset_evaluate(primitive_root, n, offset, poly));
        ASSERT_EQ(poly.evaluate_domain(domain), fast_coset_evaluate(primitive_root, n, offset, poly));
    });
    run("fft::ntt_arithmetics::tests::coset_divide", [] {                 // ntt_arithmetics.rs:494-517
        Field field(FIELD_PRIME);
        const u128 n = 1 << 6;
        FieldElement primitive_root = field.primitive_nth_root(n);
        for (int trial = 0; trial < 20; trial++) {
            Polynomial lhs = rand_poly(field, n / 2), rhs = rand_poly(field, n / 2);
            Polynomial prod = fast_multiply(primitive_root, n, lhs, rhs);
            // the quotient keeps deg(prod) - deg(lhs) + 1 coefficients: trailing zero coefficients of rhs are dropped by
            // the reference too (it compares against a rhs whose top coefficient is non-zero with overwhelming probability)
            Polynomial div = fast_coset_divide(primitive_root, n, field.generator(), prod, lhs);
            ASSERT(div == rhs, "#%d trial failed", trial);
        }
    });
    run("merkle_root::tests::commit_one", [] {                            // merkle_root.rs:107-128
        Field field(FIELD_PRIME);
        ASSERT_EQ(MerkleRoot::commit({fe(field, 11)}),
                  Bytes("7aa7e388f8145d395ac616bb526eaa35b10069f49e2b36d7327157d1d4af360dfbbfea805aa7e405ed025ce5eadd56c27c40b92991727a5a16b51df5604ad006"));
        ASSERT_EQ(MerkleRoot::commit({fe(field, 5462)}).to_hex(),
                  std::string("1f069c52b4f26c7714dbd9babacbff542d1333190e3246dec47ee9f30bb649046406f3e0ae8f4cafd52bc1a1305061b451a8746ad3ad240c2524a82a3fcd28c0"));
    });
    run("merkle_root::tests::commit_two", [] {                            // merkle_root.rs:130-157
        Field field(FIELD_PRIME);
        ASSERT_EQ(MerkleRoot::commit(fes(field, {5462ULL, 456ULL})),
                  Bytes("e79bb3f920912c56d27de11b3aaedf523d75877d7ec34d7b5819142ba69ce421e665b176fbbbd7b81e90dce61b1f629830eec87c3f7d0644c412af12f47548fe"));
        ASSERT_EQ(MerkleRoot::commit(fes(field, {652ULL, 23409ULL})),
                  Bytes("9b70e42c4b3aea3efddaeda6c1883b38c8969e40ca17566d612156c0457961e7c30d811e2adefd941da7b5329d24ecf015dcffb3e39e379dc988564d588a2341"));
    });
    run("merkle_root::tests::commit_four", [] {                           // merkle_root.rs:159-180
        Field field(FIELD_PRIME);
        ASSERT_EQ(MerkleRoot::commit(fes(field, {5462ULL, 456ULL, 652ULL, 23409ULL})),
                  Bytes("b36f5edab7ea2100fc298d9811bf1a745745282e80243e3a919e71ef6c30f690606b445557ad7843d3251c8e92b83b584d94b738334ffa7d88babd6e47471ac5"));
        bool panicked = false;
        try { MerkleRoot::commit(fes(field, {1ULL, 2ULL, 3ULL})); } catch (const Panic& p) { panicked = p.code == ZKB_ERR_NOT_POW2; }
        ASSERT(panicked, "length must be power of two (merkle_root.rs:9)");
    });
    run("merkle_root::tests::open", [] {                                  // merkle_root.rs:182-202
        Field field(FIELD_PRIME);
        auto leafs = fes(field, {5462ULL, 456ULL, 652ULL, 23409ULL});
        std::vector<Bytes> path = MerkleRoot::open(1, leafs);
        ASSERT_EQ(path, (std::vector<Bytes>{
            "1f069c52b4f26c7714dbd9babacbff542d1333190e3246dec47ee9f30bb649046406f3e0ae8f4cafd52bc1a1305061b451a8746ad3ad240c2524a82a3fcd28c0",
            "9b70e42c4b3aea3efddaeda6c1883b38c8969e40ca17566d612156c0457961e7c30d811e2adefd941da7b5329d24ecf015dcffb3e39e379dc988564d588a2341"}));
        Bytes root = MerkleRoot::commit(leafs);
        for (size_t i = 0; i < leafs.size(); i++) ASSERT(MerkleRoot::verify(root, i, MerkleRoot::open(i, leafs), leafs[i]), "opening %zu must verify", i);
    });
    auto fri_verify = [](bool native_stream) {                            // fri.rs:450-531
        Field field(FIELD_PRIME);
        const size_t degree = 63, expansion_factor = 4, num_colinearity_tests = 17;
        const size_t codeword_initial_length = (degree + 1) * expansion_factor;
        FieldElement omega = field.primitive_nth_root(codeword_initial_length);
        FieldElement generator = field.generator();
        FRI fri(generator, omega, codeword_initial_length, expansion_factor, num_colinearity_tests);
        std::vector<FieldElement> coeffs;
        for (size_t i = 0; i <= degree; i++) coeffs.push_back(fe(field, i));
        Polynomial polynomial(coeffs);
        std::vector<FieldElement> domain;
        for (size_t i = 0; i < codeword_initial_length; i++) domain.push_back(omega ^ (u128)i);
        std::vector<FieldElement> codeword = polynomial.evaluate_domain(domain);

        // a ProofStream the library knows nothing about: Fiat-Shamir goes through the per-round callback
        struct ForeignStream : ProofStream {
            IndependentProofStream inner;
            Bytes digest() const override { return inner.digest(); }
            Bytes fiat_shamir_prover(size_t n) const override { return inner.fiat_shamir_prover(n); }
            Bytes fiat_shamir_verifier(size_t n) const override { return inner.fiat_shamir_verifier(n); }
            void push(const StarkProofStreamEnum& o) override { inner.push(o); }
            std::optional<StarkProofStreamEnum> pull() override { return inner.pull(); }
        };
        auto prove_verify = [&](const std::vector<FieldElement>& cw, std::vector<std::pair<size_t, FieldElement>>& points, Bytes* proof) {
            if (native_stream) {
                IndependentProofStream ps;
                fri.prove(cw, ps);
                if (proof) *proof = ps.digest();
                return fri.verify(ps, points);
            }
            ForeignStream ps;
            fri.prove(cw, ps);
            if (proof) *proof = ps.digest();
            return fri.verify(ps, points);
        };
        std::vector<std::pair<size_t, FieldElement>> points;
        Bytes proof;
        Result res = prove_verify(codeword, points, &proof);
        ASSERT(res == Result::Ok(), "proof should be valid: %s", res.err ? res.err->c_str() : "");
        ASSERT_EQ(points.size(), 2 * num_colinearity_tests);
        for (auto& xy : points) ASSERT(polynomial.evaluate(omega ^ (u128)xy.first) == xy.second, "polynomial evaluates to wrong value");
        static Bytes first_proof;
        if (first_proof.buf.empty()) first_proof = proof;
        ASSERT(proof == first_proof, "native-stream and callback-stream proofs must be byte-identical");
        std::printf("    fri::tests::verify proof: %zu bytes, blake2b512 = %s\n", proof.buf.size(), blake2b512(proof).to_hex().c_str());
        // disturb then test for failure
        for (size_t i = 0; i < degree / 3; i++) codeword[i] = field.zero();
        points.clear();
        ASSERT(prove_verify(codeword, points, nullptr) != Result::Ok(), "proof should fail, but is accepted");
    };
    run("fri::tests::verify (library proof stream: one zkb_fri_prove call)", [&] { fri_verify(true); });
    run("fri::tests::verify (foreign ProofStream: Fiat-Shamir callback per round)", [&] { fri_verify(false); });
    run("fri: verify rejects tampered proofs with the reference's messages", [] {     // fri.rs:279-281, 374-383, 392-408
        Field field(FIELD_PRIME);
        const size_t n = 256, ef = 4, ncc = 17;
        FieldElement omega = field.primitive_nth_root(n);
        FRI fri(field.generator(), omega, n, ef, ncc);
        std::vector<FieldElement> coeffs, domain;
        for (size_t i = 0; i < n / ef; i++) coeffs.push_back(fe(field, 3 * i + 1));
        std::vector<FieldElement> codeword = fast_coset_evaluate(omega, n, field.one(), Polynomial(coeffs));
        IndependentProofStream honest;
        fri.prove(codeword, honest);
        const size_t R = fri.num_rounds();
        auto verdict = [&](const std::function<void(std::vector<StarkProofStreamEnum>&)>& tamper) {
            std::vector<StarkProofStreamEnum> objs = honest.objects;
            tamper(objs);
            IndependentProofStream ps(objs);
            std::vector<std::pair<size_t, FieldElement>> points;
            Result r = fri.verify(ps, points);
            return r.err ? *r.err : std::string("ok");
        };
        ASSERT_EQ(verdict([](std::vector<StarkProofStreamEnum>&) {}), std::string("ok"));
        ASSERT_EQ(objs_kind(honest.objects, R), (int)StarkProofStreamEnum::Codeword);
        // the last codeword no longer matches the last root
        ASSERT_EQ(verdict([&](std::vector<StarkProofStreamEnum>& o) { o[R].codeword[3] = o[R].codeword[3] + field.one(); }),
                  std::string("last codeword is not well formed"));
        // a first-layer leaf changed: the challenges stay the same (Leafs come after every challenge), the colinearity test fails
        ASSERT_EQ(verdict([&](std::vector<StarkProofStreamEnum>& o) { o[R + 1].leafs[2] = o[R + 1].leafs[2] + field.one(); }),
                  std::string("colinearity check failure"));
        // a path node changed: Merkle verification of the first opened leaf fails
        ASSERT_EQ(verdict([&](std::vector<StarkProofStreamEnum>& o) { o[R + 1 + ncc].path[0].buf[5] ^= 1; }),
                  std::string("Merkle auth path verification failed for aa"));
        ASSERT_EQ(verdict([&](std::vector<StarkProofStreamEnum>& o) { o[R + 1 + ncc + 2].path[1].buf[0] ^= 0x80; }),
                  std::string("Merkle auth path verification failed for cc"));
        // a root changed: every later challenge changes, so the proof cannot verify (which check trips first depends on the challenge)
        ASSERT_NE(verdict([&](std::vector<StarkProofStreamEnum>& o) { o[0].root.buf[0] ^= 1; }), std::string("ok"));
        // an object of the wrong kind where a root is expected panics (expect_root, proof_stream_enum.rs:129-133)
        bool panicked = false;
        try { verdict([&](std::vector<StarkProofStreamEnum>& o) { o[1] = StarkProofStreamEnum::Value_(field.one()); }); } catch (const Panic&) { panicked = true; }
        ASSERT(panicked, "expect_root on a Value must panic");
    });
    run("fri: prove panics on a codeword of the wrong length", [] {        // fri.rs:215-219
        Field field(FIELD_PRIME);
        FRI fri(field.generator(), field.primitive_nth_root(256), 256, 4, 17);
        IndependentProofStream ps;
        bool panicked = false;
        try { fri.prove(std::vector<FieldElement>(128, field.one()), ps); } catch (const Panic& p) {
            panicked = std::string(p.what()) == "Length of the domain doesnt match the length of initial codeword";
        }
        ASSERT(panicked, "length mismatch must panic");
    });
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "all";
    host_tests();
    if (mode != "host") gpu_tests();
    std::printf("\ntest result: %s. %d passed; %d failed (%s)\n", g_fail ? "FAILED" : "ok", g_run - g_fail, g_fail, mode.c_str());
    return g_fail ? 1 : 0;
}
