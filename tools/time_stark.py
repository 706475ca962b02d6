"""Where the time goes in zk.Stark.prove (the device prover) for one RPSSS signature: wall-clock per stage.
Usage (on the GPU box): python tools/time_stark.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_stark_tutor_b200 as zk                                   # noqa: E402
from zk_stark_tutor_b200 import fft, stark as S                     # noqa: E402
from oracle.stark import RPSSS, deterministic_rng                  # noqa: E402

ctx = zk.Context(0)
cpu = RPSSS(4, 64, 128, 3)
sk, pk = cpu.keygen(deterministic_rng(b"k2"))
tcs = [tc.dictionary for tc in cpu.transition_constraints()]
st = zk.Stark(4, 64, 128, cpu.rp.m, cpu.rp.N + 1, 3, ctx=ctx)
trace, boundary = cpu.rp.trace(sk), cpu.rp.boundary_constraints(pk)

stages = {}


def timed(name, fn):
    def wrap(*a, **k):
        ctx.sync()
        t = time.perf_counter()
        r = fn(*a, **k)
        ctx.sync()
        stages[name] = stages.get(name, 0.0) + time.perf_counter() - t
        return r
    return wrap


fft.fast_interpolate_domain = timed("fast_interpolate_domain", fft.fast_interpolate_domain)
fft.fast_zerofier = timed("fast_zerofier", fft.fast_zerofier)
fft.fast_coset_divide = timed("fast_coset_divide", fft.fast_coset_divide)
S.air_combination = timed("air_combination", S.air_combination)
st._lde_into = timed("lde", st._lde_into)
st._coset_degree = timed("degree_check", st._coset_degree)
st.fri.prove = timed("fri_prove", st.fri.prove)
S.MerkleTree.open_into = timed("open_into", S.MerkleTree.open_into)

for rep in range(3):
    stages.clear()
    t0 = time.perf_counter()
    sig = st.prove(trace, tcs, boundary, zk.SignatureProofStream(b"doc"), deterministic_rng(b"r"), lockstep=False)
    total = time.perf_counter() - t0
    print("rep %d: total %.1f ms, %d bytes" % (rep, total * 1e3, len(sig)))
    for k, v in sorted(stages.items(), key=lambda kv: -kv[1]):
        print("    %-26s %8.2f ms" % (k, v * 1e3))
    print("    %-26s %8.2f ms" % ("(other host work)", (total - sum(stages.values())) * 1e3))

# the same signature through the batched entry points with a batch of one (every stage a single device call)
import time as _t                                                   # noqa: E402
for rep in range(3):
    ctx.sync()
    t0 = _t.perf_counter()
    out = st.prove_batch([trace], tcs, [boundary], [zk.SignatureProofStream(b"doc")], [deterministic_rng(b"r")])
    print("prove_batch, batch of 1: %.2f ms, %d bytes, same bytes as prove: %s" % ((_t.perf_counter() - t0) * 1e3, len(out[0]), out[0] == sig))
