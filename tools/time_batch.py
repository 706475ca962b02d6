"""Where the time goes in Stark.prove_batch (batch of 32 real RPSSS signatures, ONE lane): wall-clock per C call vs Python glue.
Usage (on the GPU box): python tools/time_batch.py [batch]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                                 # noqa: E402
import zk_stark_tutor_b200 as zk                                   # noqa: E402
from zk_stark_tutor_b200.context import pack                       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fx = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rpsss_air.json")))
pr = fx["params"]
ctx = zk.Context(0)
st = zk.Stark(pr["expansion_factor"], pr["num_collinearity_checks"], pr["security_level"], pr["num_registers"], pr["num_cycles"],
              pr["transition_constraints_degree"], ctx=ctx)
tcs = [{tuple(k): int(v) for k, v in tc} for tc in fx["transition_constraints"]]
cases = fx["cases"]
traces = [pack([int(v) for row in c["trace"] for v in row]).reshape(len(c["trace"]), pr["num_registers"], 2) for c in cases]
bounds = [[(cy, reg, int(v)) for cy, reg, v in c["boundary"]] for c in cases]

stages = {}
lib = ctx.lib


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t = time.perf_counter()
        r = self.fn(*a)
        stages[self.name] = stages.get(self.name, 0.0) + time.perf_counter() - t
        return r


class LibProxy:
    def __getattr__(self, name):
        return Timed(name, getattr(lib, name))


ctx.lib = LibProxy()
for rep in range(4):
    stages.clear()
    idx = [i % len(cases) for i in range(B)]
    t0 = time.perf_counter()
    st.prove_batch([traces[i] for i in idx], tcs, [bounds[i] for i in idx], [zk.SignatureProofStream(cases[i]["document"].encode()) for i in idx],
                   [os.urandom] * B, return_bytes=False)
    total = time.perf_counter() - t0
    c_time = sum(stages.values())
    print("rep %d: batch %d: %.2f ms total = %.1f us / signature; C calls %.2f ms, Python glue %.2f ms (%.1f us / signature)"
          % (rep, B, total * 1e3, total / B * 1e6, c_time * 1e3, (total - c_time) * 1e3, (total - c_time) / B * 1e6))
    if rep == 3:
        for k, v in sorted(stages.items(), key=lambda kv: -kv[1])[:12]:
            print("    %-32s %8.3f ms" % (k, v * 1e3))

import cProfile                                                    # noqa: E402
import pstats                                                      # noqa: E402
ctx.lib = lib
idx = [i % len(cases) for i in range(B)]
prof = cProfile.Profile()
prof.enable()
for _ in range(5):
    st.prove_batch([traces[i] for i in idx], tcs, [bounds[i] for i in idx], [zk.SignatureProofStream(cases[i]["document"].encode()) for i in idx],
                   [os.urandom] * B, return_bytes=False)
prof.disable()
pstats.Stats(prof).sort_stats("tottime").print_stats(18)
