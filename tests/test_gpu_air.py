"""SURVEY.md 8f.3 / 8f.4 on the B200: the transition quotients, x^shift products and the weighted combination of
Stark::prove (stark.rs:388-519) in evaluation form (zkb_air_combination) against the oracle's coefficient-form
route, and the whole Stark::prove / RPSSS signature through zk.Stark - proof bytes identical to the oracle's."""
import numpy as np
import pytest

import zk_stark_tutor_b200 as zk
from air_common import prover_intermediates
from oracle import cbind as C, ntt as N, proof_stream as PS
from oracle.stark import RPSSS, deterministic_rng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def cuda(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def host(t):
    return C.from_arr(t.cpu().numpy().view(np.uint64))


def test_air_combination_equals_coefficient_form_prover(ctx):
    m = prover_intermediates()
    st, nr, nc, n = m["stark"], m["nr"], m["nc"], m["n"]
    bq = cuda(np.stack([C.to_arr(cw) for cw in m["bq_cws"]]))
    rnd = cuda(C.to_arr(m["rnd_cw"]))
    before = ctx.launches
    comb, tq = zk.air_combination(st.generator, st.omega, n, st.expansion_factor, m["tcs"], m["zerofiers"], m["interpolants"], m["tz"],
                                  m["weights"], m["shifts"], bq, rnd, want_quotients=True, ctx=ctx)
    assert ctx.launches > before
    assert host(comb) == m["combined"], "evaluation-form combination differs from LDE(weighted sum of terms)"
    for j in range(nc):
        assert host(tq[j]) == N.fast_coset_evaluate(st.omega, n, st.generator, m["tq_polys"][j]), "transition quotient %d" % j
    # without the quotient output
    comb2 = zk.air_combination(st.generator, st.omega, n, st.expansion_factor, m["tcs"], m["zerofiers"], m["interpolants"], m["tz"],
                               m["weights"], m["shifts"], bq, rnd, ctx=ctx)
    assert host(comb2) == m["combined"]


def test_air_combination_rejects_host_codewords_and_zero_divisor(ctx):
    import ctypes
    from zk_stark_tutor_b200 import _lib
    m = prover_intermediates()
    st, nr, n = m["stark"], m["nr"], m["n"]
    bq = cuda(np.stack([C.to_arr(cw) for cw in m["bq_cws"]]))
    rnd = cuda(C.to_arr(m["rnd_cw"]))
    # a transition zerofier with a root ON the coset: x - generator  ->  the reference's division panics
    bad_tz = [(-st.generator) % zk.P, 1]
    with pytest.raises(zk.ZkbError) as e:
        zk.air_combination(st.generator, st.omega, n, st.expansion_factor, m["tcs"], m["zerofiers"], m["interpolants"], bad_tz,
                           m["weights"], m["shifts"], bq, rnd, ctx=ctx)
    assert e.value.code == -7
    d = _lib.AirDesc()
    assert ctx.lib.zkb_air_combination(ctx.h, ctypes.byref(d), None, 0, None, None, None) == -2


def test_rpsss_signature_through_device_stark(ctx):
    """Stark::prove on the Rescue-Prime trace at the tutorial's signature parameters with the evaluation-form middle:
    the 1,156,888-byte signature equals the oracle's (coefficient-form) one and the restated verifier accepts it."""
    cpu = RPSSS(4, 64, 128, 3)
    sk, pk = cpu.keygen(deterministic_rng(b"k2"))
    doc = b"evaluation form"
    want = cpu.sign(sk, doc, deterministic_rng(b"r2"))
    stark = zk.Stark(4, 64, 128, cpu.rp.m, cpu.rp.N + 1, 3, ctx=ctx)
    assert (stark.omicron_domain_length, stark.fri_domain_length) == (1024, 4096)
    tcs = [tc.dictionary for tc in cpu.transition_constraints()]
    got = stark.prove(cpu.rp.trace(sk), tcs, cpu.rp.boundary_constraints(pk), zk.SignatureProofStream(doc), deterministic_rng(b"r2"))
    assert len(got) == 1156888
    assert got == want
    # the reference's call structure (one call per polynomial) gives the same bytes as the lockstep pipeline
    assert stark.prove(cpu.rp.trace(sk), tcs, cpu.rp.boundary_constraints(pk), zk.SignatureProofStream(doc), deterministic_rng(b"r2"),
                       lockstep=False) == want
    assert cpu.verify(pk, doc, got) is None
    assert cpu.verify(pk, b"another document", got) is not None
    # the one-call batched prover (zkb_stark_prove_batch) drawing its randomizers from OS entropy: a fresh proof that the restated verifier accepts
    import os
    fresh = stark.prove_batch([cpu.rp.trace(sk)], tcs, [cpu.rp.boundary_constraints(pk)], [zk.SignatureProofStream(doc)], [os.urandom])[0]
    assert len(fresh) == 1156888 and fresh != got
    assert cpu.verify(pk, doc, fresh) is None
    # a trace that violates the AIR is caught by the degree check (stark.rs:451-464)
    bad = [list(r) for r in cpu.rp.trace(sk)]
    bad[5][1] = (bad[5][1] + 1) % zk.P
    with pytest.raises(ValueError):
        stark.prove(bad, tcs, cpu.rp.boundary_constraints(pk), zk.SignatureProofStream(doc), deterministic_rng(b"r2"))


def test_stark_prove_independent_stream_through_device_stark(ctx):
    """stark.rs:810-880: hash-trace proof over an IndependentProofStream (BASELINE configs[0]), device prover == oracle."""
    from oracle.rescue_prime import RescuePrime
    from oracle.stark import Stark
    rp = RescuePrime(2, 1, 128, 27)
    ref = Stark(4, 64, 128, rp.m, rp.N + 1, 3)
    x = 0xC0FFEE1234
    out = rp.hash(x)
    tcs = rp.transition_constraints(ref.omicron, ref.omicron_domain_length)
    want = ref.prove(rp.trace(x), tcs, rp.boundary_constraints(out), PS.IndependentProofStream(), deterministic_rng(b"q"))
    dev = zk.Stark(4, 64, 128, rp.m, rp.N + 1, 3, ctx=ctx)
    got = dev.prove(rp.trace(x), tcs, rp.boundary_constraints(out), zk.IndependentProofStream(), deterministic_rng(b"q"), check_degrees=False)
    assert got == want
    assert dev.prove(rp.trace(x), tcs, rp.boundary_constraints(out), zk.IndependentProofStream(), deterministic_rng(b"q"), lockstep=False) == want
    assert ref.verify(tcs, rp.boundary_constraints(out), PS.IndependentProofStream(PS.parse(got))) is None


def test_committed_rpsss_fixture_signatures(ctx):
    """tests/golden/rpsss_air.json (AIR, traces, boundary conditions as data + the oracle prover's signature digests):
    zk.Stark.prove reproduces every committed digest - the check bench.py's `signatures` workload repeats before timing."""
    import hashlib
    import json
    import os
    from zk_stark_tutor_b200.stark import deterministic_rng as drng
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpsss_air.json")))
    pr = fx["params"]
    stark = zk.Stark(pr["expansion_factor"], pr["num_collinearity_checks"], pr["security_level"], pr["num_registers"], pr["num_cycles"],
                     pr["transition_constraints_degree"], ctx=ctx)
    tcs = [{tuple(k): int(v) for k, v in tc} for tc in fx["transition_constraints"]]
    for k, c in enumerate(fx["cases"]):
        sig = stark.prove([[int(v) for v in row] for row in c["trace"]], tcs, [(cy, reg, int(v)) for cy, reg, v in c["boundary"]],
                          zk.SignatureProofStream(c["document"].encode()), drng(c["rng_seed"].encode()), lockstep=bool(k % 2))
        assert len(sig) == c["signature_bytes"] == 1156888
        assert hashlib.sha256(sig).hexdigest() == c["signature_sha256"]


def _fixture():
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpsss_air.json")))
    tcs = [{tuple(k): int(v) for k, v in tc} for tc in fx["transition_constraints"]]
    cases = [dict(trace=[[int(v) for v in row] for row in c["trace"]], boundary=[(cy, reg, int(v)) for cy, reg, v in c["boundary"]],
                  doc=c["document"].encode(), seed=c["rng_seed"].encode(), sha=c["signature_sha256"]) for c in fx["cases"]]
    return fx["params"], tcs, cases


def test_trace_lde_batch_equals_interpolate_then_lde(ctx):
    """zkb_trace_lde_batch (iNTT mod the prefix zerofier, then coset LDE) == fast_interpolate_domain + fast_coset_evaluate (stark.rs:305-326, 373-378)"""
    import ctypes
    import random
    import torch
    from oracle import field as F, poly as PL
    from zk_stark_tutor_b200.context import le16
    rnd = random.Random(77)
    for L, N_, n in ((284, 1024, 4096), (5, 8, 32), (8, 8, 16), (2, 4, 16), (100, 128, 512)):
        omicron, omega = F.primitive_nth_root(N_), F.primitive_nth_root(n)
        dom = [F.fpow(omicron, i) for i in range(L)]
        cols = [[rnd.randrange(zk.P) for _ in range(L)] for _ in range(3)] + [[0] * L, [5] * L]
        vals = np.stack([C.to_arr(c) for c in cols])
        out = torch.empty((len(cols), n, 2), dtype=torch.int64, device="cuda")
        coeffs = torch.empty((len(cols), L, 2), dtype=torch.int64, device="cuda")
        ctx.check(ctx.lib.zkb_trace_lde_batch(ctx.h, le16(omicron), N_, L, le16(omega), n, le16(F.GENERATOR), vals.ctypes.data, L, len(cols),
                                              out.data_ptr(), n, coeffs.data_ptr()))
        for k, col in enumerate(cols):
            want = PL.fast_interpolate_domain(omicron, N_, dom, col)
            d = PL.degree(want)
            got = host(coeffs[k])
            assert PL.degree(got) == d and (d is None or got[:d + 1] == want[:d + 1]), (L, N_, k)
            assert host(out[k]) == N.fast_coset_evaluate(omega, n, F.GENERATOR, got), (L, N_, k)
        degs = (ctypes.c_int64 * len(cols))()
        ctx.check(ctx.lib.zkb_coset_degree_batch(ctx.h, le16(omega), out.data_ptr(), n, n, len(cols), degs))
        assert list(degs) == [(-1 if PL.degree(host(coeffs[k])) is None else PL.degree(host(coeffs[k]))) for k in range(len(cols))]


def test_prove_batch_reproduces_committed_signatures(ctx):
    """Stark.prove_batch: four different signatures (keys, documents, randomness) advancing in lockstep == the committed digests of
    the oracle's one-at-a-time coefficient-form prover; a batch of one as well."""
    import hashlib
    from zk_stark_tutor_b200.stark import deterministic_rng as drng
    pr, tcs, cases = _fixture()
    stark = zk.Stark(pr["expansion_factor"], pr["num_collinearity_checks"], pr["security_level"], pr["num_registers"], pr["num_cycles"],
                     pr["transition_constraints_degree"], ctx=ctx)
    before = ctx.launches
    sigs = stark.prove_batch([c["trace"] for c in cases], tcs, [c["boundary"] for c in cases], [zk.SignatureProofStream(c["doc"]) for c in cases],
                             [drng(c["seed"]) for c in cases])
    launches = ctx.launches - before
    assert [len(x) for x in sigs] == [1156888] * len(cases)
    assert [hashlib.sha256(x).hexdigest() for x in sigs] == [c["sha"] for c in cases]
    c = cases[2]
    one = stark.prove_batch([c["trace"]], tcs, [c["boundary"]], [zk.SignatureProofStream(c["doc"])], [drng(c["seed"])], check_degrees=False)
    assert hashlib.sha256(one[0]).hexdigest() == c["sha"]
    assert ctx.launches - before - launches <= launches              # the launch count does not grow with the batch
    # a trace that violates the AIR is caught by the degree check (stark.rs:451-464)
    bad = [list(r) for r in c["trace"]]
    bad[7][0] = (bad[7][0] + 1) % zk.P
    with pytest.raises(ValueError):
        stark.prove_batch([c["trace"], bad], tcs, [c["boundary"]] * 2, [zk.SignatureProofStream(c["doc"]) for _ in range(2)], [drng(b"a"), drng(b"b")])
    # the stage-by-stage Python call sequence (native=False) and the single C call (zkb_stark_prove_batch, the default) give the same bytes
    staged = stark.prove_batch([c["trace"] for c in cases], tcs, [c["boundary"] for c in cases], [zk.SignatureProofStream(c["doc"]) for c in cases],
                               [drng(c["seed"]) for c in cases], native=False)
    assert staged == sigs
    with pytest.raises(ValueError):
        stark.prove_batch([c["trace"], bad], tcs, [c["boundary"]] * 2, [zk.SignatureProofStream(c["doc"]) for _ in range(2)], [drng(b"a"), drng(b"b")],
                          native=False)
    # OS entropy (randomness = NULL): a different, valid proof of the same size every time; packed-array traces
    import os
    packed = zk.stark.pack([v for row in c["trace"] for v in row]).reshape(len(c["trace"]), pr["num_registers"], 2)
    a = stark.prove_batch([packed] * 2, tcs, [c["boundary"]] * 2, [zk.SignatureProofStream(c["doc"]) for _ in range(2)], [os.urandom] * 2)
    assert [len(x) for x in a] == [1156888] * 2 and a[0] != a[1]
    stark.close()
