"""The oracle against every known-answer test the reference holds for the hot path.

Each block cites the reference test it re-asserts.  CPU only.
"""
from oracle import field as F
from oracle import ntt as N
from oracle import merkle as M
from oracle import proof_stream as PS
from oracle.fri import FRI, test_colinearity as colinear

P = F.P


def test_field_constants():
    assert P == 270497897142230380135924736767050121217          # field.rs:9-10
    assert P == 0xCB800000_00000000_00000000_00000001
    assert F.fpow(F.GENERATOR, 1 << 119) == 1 and F.fpow(F.GENERATOR, 1 << 118) != 1


def test_field_element_kats():
    # src/field/field_element.rs:151-299
    assert F.mul(49789714223038013592473676705012096123, 6534789852937546098347957826345234) \
        == 105250150227149389100670877502232671566
    assert F.mul(8, 12) == 96
    assert F.mul(3, P - 2) == P - 6
    assert F.div(74658620945386735627456854792784352353, F.GENERATOR) \
        == 120557879365253444230411244907275635216
    assert F.div(12, 4) == 3
    assert F.div(P - 2, 5) == 54099579428446076027184947353410024243
    assert F.div(5012096123, 6534789852937546098347957826345234) \
        == 109071144973379706934869779239844248849
    assert F.inv(256) == 269441264731518542713518780764053831681
    assert F.mul(8, F.inv(8)) == 1 and F.mul(P - 2, F.inv(P - 2)) == 1
    assert F.add(270497897142230380135924736767050120961, 300) == 44
    assert F.sub(44, 200) == 270497897142230380135924736767050121061
    assert F.neg(6534789852937546098) == 270497897142230380129389946914112575119
    assert F.fpow(6534789852937546098, 501209126122) == 256557788041265930815463337858691703671
    assert F.fpow(15, 4) == 50625
    assert F.fpow(270497897142230380135, 8) == 79016866124691016201920330826259043252


def test_field_kats():
    # src/field/field.rs:177-256
    assert F.mul(2, 3) == 6 and F.mul(P - 1, 3) == P - 3
    assert F.primitive_nth_root(256) == 178902808384765167578311106676137348214
    assert F.primitive_nth_root(2) == P - 1
    z = F.primitive_nth_root(256)
    assert F.fpow(z, 256) == 1 and F.fpow(z, 128) != 1
    assert F.sample(bytes.fromhex("6c9c4992")) == 1822181778
    assert F.sample(bytes.fromhex("ac4cd3be")) == 2890716094
    assert F.neg(256) == 270497897142230380135924736767050120961
    assert F.inv(2) == 135248948571115190067962368383525060609


NTT16_IN = [10350860596407318609598574026175964133, 60692809610834653822383343680910625982, 223446197944610152228521360138425742723, 123599176902523769876954930401435714041, 233214499950980668770362073427851594143, 197530481770421435151547222505733630031, 6028204552208455457232478170590637777, 129106051215868132791440857107220454376, 46875137253396986423834480299002499296, 40573479539486208028801437611599580111, 177627388180112816822358878396956962568, 63754231379381382860231899477157171256, 213977912421556511151382836938765186268, 247295448209556494808801789962732329479, 198078312580458497833840274756537503682, 140348661180454074099943144751461445367]
NTT16_OUT = [219013573292644897785762424283206192714, 28020178707455534238013018981848447223, 125720672179066355667363683873634014638, 9544075888957995047526079628773702483, 236214009288214032104373542256167121711, 203576991437594049347129945434757211067, 161303837601531457486430204397030363075, 8066037348193233635957451882263404827, 106698173671205255857026330656055139947, 205516443913240407551582667880265743260, 132452175458644240344865798387130681692, 14403130148933356826258737037147692544, 103258398926149853393925877736501914903, 241567637358481607032146874821122458208, 184640833807669488035490783312852642403, 79102880510147994351196921219409543952]
INTT16_VALUES = [159, 179, 197, 143, 198, 82, 100, 153, 45, 158, 154, 238, 46, 121, 148, 200]
INTT16_COEFFS = [2321, 46679697743149797158402415879589215379, 85767599764045409871854383990500128680, 170048455543476672374689900824216177289, 56517926799859326797837626323965682333, 150718635918560071455504820257610329093, 149093701728889244918633279335367822666, 266977550113122771518657412035427200127, 270497897142230380135924736767050120990, 63434915687244166391766073758524869310, 261359683971832165794823307314869483630, 172866549451408128829178953127270691728, 213979970342371053338087110443084438582, 83513730222590766426030683639871516493, 44774808819693939686538502893362807298, 127752053889369146389468687545690486361]


def test_ntt_kats():
    # src/fft/ntt.rs:78-130
    w = F.primitive_nth_root(16)
    assert N.ntt(w, NTT16_IN) == NTT16_OUT
    # cross-check vs naive evaluation like the reference test does
    naive = [sum(c * F.fpow(w, i * k) for i, c in enumerate(NTT16_IN)) % P for k in range(16)]
    assert naive == NTT16_OUT
    assert N.ntt(w, INTT16_VALUES) == INTT16_COEFFS
    assert N.intt(w, INTT16_COEFFS) == INTT16_VALUES


def test_ntt_padding_and_small():
    w8 = F.primitive_nth_root(8)
    assert N.ntt(w8, [1, 2, 3, 4, 5]) == N.ntt(w8, [1, 2, 3, 4, 5, 0, 0, 0])
    assert N.ntt(w8, [7]) == [7] and N.intt(w8, [7]) == [7]
    assert N.intt(w8, N.ntt(w8, [1, 2, 3, 4, 5])) == [1, 2, 3, 4, 5, 0, 0, 0]


def test_scale_kat():
    # src/field/polynomial.rs:632-652 scales by 4: c_i -> 4^i c_i
    assert N.scale([1, 2, 3], 4) == [1, 8, 48]


def test_fast_arithmetic_vs_schoolbook():
    # src/fft/ntt_arithmetics.rs:356-517 (property tests, n = 64)
    import random
    rnd = random.Random(7)
    n = 64
    w = F.primitive_nth_root(n)
    for _ in range(5):
        a = [rnd.randrange(P) for _ in range(rnd.randrange(1, 31))]
        b = [rnd.randrange(P) for _ in range(rnd.randrange(1, 31))]
        school = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                school[i + j] = (school[i + j] + x * y) % P
        assert N.fast_multiply(w, n, a, b) == school
        # coset evaluate == naive evaluation on offset * w^k
        off = F.GENERATOR
        ev = N.fast_coset_evaluate(w, n, off, a)
        for k in (0, 1, 17, 63):
            x = off * F.fpow(w, k) % P
            assert ev[k] == sum(c * F.fpow(x, i) for i, c in enumerate(a)) % P
        # coset divide: (a*b)/b == a
        assert N.fast_coset_divide(w, n, off, school, b) == a


def test_blake2b_kats():
    # src/crypto/blake2b512.rs:22-30
    assert M.blake2b512(b"\x00").hex() == "2fa3f686df876995167e7c2e5d74c4c7b6e48f8068fe0e44208344d480f7904c36963e44115fe3eb2a3ac8694c28bcb4f5a0f3276f2e79487d8219057a506e4b"
    assert M.blake2b512(b"\x00\x00").hex() == "5ba7f7e4ade7e5803c59d184326420823f7f860effcfba0bb896d568f59b8d85181cfff25929d40b18e01069c2ef5c31754f1d821a1f3f80f896f4dde374a2f1"


H_11 = "7aa7e388f8145d395ac616bb526eaa35b10069f49e2b36d7327157d1d4af360dfbbfea805aa7e405ed025ce5eadd56c27c40b92991727a5a16b51df5604ad006"
H_5462 = "1f069c52b4f26c7714dbd9babacbff542d1333190e3246dec47ee9f30bb649046406f3e0ae8f4cafd52bc1a1305061b451a8746ad3ad240c2524a82a3fcd28c0"
H_5462_456 = "e79bb3f920912c56d27de11b3aaedf523d75877d7ec34d7b5819142ba69ce421e665b176fbbbd7b81e90dce61b1f629830eec87c3f7d0644c412af12f47548fe"
H_652_23409 = "9b70e42c4b3aea3efddaeda6c1883b38c8969e40ca17566d612156c0457961e7c30d811e2adefd941da7b5329d24ecf015dcffb3e39e379dc988564d588a2341"
H_4 = "b36f5edab7ea2100fc298d9811bf1a745745282e80243e3a919e71ef6c30f690606b445557ad7843d3251c8e92b83b584d94b738334ffa7d88babd6e47471ac5"


def test_merkle_kats():
    # src/merkle_root.rs:107-244
    assert M.commit([11]).hex() == H_11
    assert M.commit([5462]).hex() == H_5462
    assert M.commit([5462, 456]).hex() == H_5462_456
    assert M.commit([652, 23409]).hex() == H_652_23409
    assert M.commit([5462, 456, 652, 23409]).hex() == H_4
    assert [h.hex() for h in M.open_(1, [5462, 456, 652, 23409])] == [H_5462, H_652_23409]
    path = [bytes.fromhex(H_5462), bytes.fromhex(H_652_23409)]
    root = bytes.fromhex(H_4)
    assert M.verify(root, 1, path, 456)
    assert not M.verify(root, 1, path, 5462)
    assert not M.verify(root, 0, path, 456)


def test_sample_indices_kat():
    # src/fri.rs:426-448
    seed = bytes.fromhex("d4b6e8af1114859c1c24b6496a3aef2f55a21105bc103af7e12dc3b2c101fe66")
    assert FRI.sample_indices(seed, 128, 128, 17) == \
        [40, 121, 5, 113, 97, 68, 126, 88, 26, 82, 81, 91, 93, 125, 10, 57, 48]


def test_shake_kats():
    # src/proof_stream.rs:88-146 (the stream there digests as its Debug string)
    assert PS.shake256(b"[]", 64).hex() == "ec784925b52067bce01fd820f554a34a3f8522b337f82e00ea03d3fa2b207ef9c2c1b9ed900cf2bbfcd19a232a94c6121e041615305c4155d46d52f58a8cff1c"
    s = b'[Str("Hello, World!"), Vec([0, 1, 5, 234]), Map({"something": 123})]'
    assert PS.shake256(s, 4).hex() == "78b0db5c"
    assert PS.shake256(s, 64).hex() == "78b0db5cfd13c78498fd0951a9fd609f2521fd02d850cc561eced844bb0c338588358abcc0d98d76c6779cb388514f4bc19e2c0125b143abee166cb98c38a831"


def test_proof_stream_roundtrip():
    # src/stark/stark.rs:785-808
    objs = [(PS.ROOT, bytes([0x49, 0x6e, 0x20, 0x74])),
            (PS.CODEWORD, [20, 100]),
            (PS.PATH, [bytes([0x49, 0x6e, 0x20, 0x74]), bytes([0x1, 0x6b, 0xfe, 0x25])]),
            (PS.LEAFS, (1, 5, 10)),
            (PS.VALUE, 2)]
    d = PS.digest(objs)
    assert d[:16] == P.to_bytes(16, "big")
    assert PS.parse(d) == objs
    assert PS.digest([(PS.ROOT, b"\x00" * 64)])[:16] == bytes(16)       # Root-only => zero header


def test_proof_size_formula():
    # src/rpsss.rs:89: 1,156,888 bytes at (ef 4, ncc 64, FRI domain 4096)
    ncc, n = 64, 4096
    root = 9 + 64
    path = lambda depth: 9 + depth * 72
    total = 16 + 3 * root                                     # 2 registers + randomizer roots
    rounds = FRI(F.GENERATOR, F.primitive_nth_root(n), n, 4, ncc).num_rounds()
    assert rounds == 4
    total += rounds * root + (9 + 512 * 16)                   # FRI roots + last codeword
    for r in range(rounds - 1):
        depth = (n >> r).bit_length() - 1
        total += ncc * (9 + 48) + ncc * (2 * path(depth) + path(depth - 1))
    total += 3 * 4 * ncc * ((9 + 16) + path(12))              # Stark openings: 3 codewords x 256 idx
    assert total == 1156888


def test_fri_roundtrip_like_reference():
    # src/fri.rs:451-531: degree-63 polynomial, N = 256, ef 4, 17 colinearity tests
    degree, ef, ncc = 63, 4, 17
    n = (degree + 1) * ef
    w = F.primitive_nth_root(n)
    fri = FRI(F.GENERATOR, w, n, ef, ncc)
    poly = list(range(degree + 1))
    codeword = N.ntt(w, poly + [0] * (n - len(poly)))
    ps = PS.IndependentProofStream()
    fri.prove(codeword, ps)
    points = []
    assert fri.verify(ps, points) is None
    for x, y in points:
        assert sum(c * F.fpow(w, x * i) for i, c in enumerate(poly)) % P == y
    bad = [0] * (degree // 3) + codeword[degree // 3:]
    ps = PS.IndependentProofStream()
    fri.prove(bad, ps)
    assert fri.verify(ps, []) is not None
    # serialised proof survives a parse round trip and still verifies
    ps = PS.IndependentProofStream()
    fri.prove(codeword, ps)
    ps2 = PS.IndependentProofStream(PS.parse(ps.digest()))
    assert fri.verify(ps2, []) is None


def test_colinearity():
    assert colinear([(1, 3), (2, 5), (5, 11)])
    assert not colinear([(1, 3), (2, 5), (5, 12)])
    assert not colinear([(1, 3), (2, 3), (5, 3)])      # degree 0 is rejected (polynomial.rs:172)


def test_domain_algorithms_vs_schoolbook():
    # src/fft/ntt_arithmetics.rs:382-468: fast_zerofier / fast_evaluate_domain /
    # fast_interpolate_domain against their schoolbook counterparts on random domains
    import random
    from oracle import poly as PL
    rnd = random.Random(11)
    n = 64
    w = F.primitive_nth_root(n)
    for size in (1, 2, 5, 16, 23):
        domain = [rnd.randrange(P) for _ in range(size)]
        assert PL.fast_zerofier(w, n, domain) == PL.zerofier_domain(domain)
        poly = [rnd.randrange(P) for _ in range(rnd.randrange(1, 31))]
        assert PL.fast_evaluate_domain(w, n, poly, domain) == [PL.evaluate(poly, x) for x in domain]
        values = [rnd.randrange(P) for _ in range(size)]
        interp = PL.fast_interpolate_domain(w, n, domain, values)
        assert [PL.evaluate(interp, x) for x in domain] == values
        assert (N.degree(interp) or 0) < size
    assert PL.fast_zerofier(w, n, []) == [] and PL.fast_interpolate_domain(w, n, [], []) == []
    q, r = PL.divide_with_rem([1, 2, 3, 4], [1, 1])
    assert PL.add(PL.mul(q, [1, 1]), r)[:4] == [1, 2, 3, 4]
