"""One lockstep batch of real RPSSS signatures (for ncu launch lists): python tools/run_signatures_once.py [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zk_stark_tutor_b200 as zk                                   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "rpsss_air.json")))
pr = fx["params"]
ctx = zk.Context(0)
stark = zk.Stark(pr["expansion_factor"], pr["num_collinearity_checks"], pr["security_level"], pr["num_registers"], pr["num_cycles"],
                 pr["transition_constraints_degree"], ctx=ctx)
tcs = [{tuple(k): int(v) for k, v in tc} for tc in fx["transition_constraints"]]
cases = [fx["cases"][i % len(fx["cases"])] for i in range(B)]
for rep in range(2):                                               # the first pass builds the cached tables
    l0 = ctx.launches
    sizes = stark.prove_batch([[[int(v) for v in row] for row in c["trace"]] for c in cases], tcs,
                              [[(cy, reg, int(v)) for cy, reg, v in c["boundary"]] for c in cases],
                              [zk.SignatureProofStream(c["document"].encode()) for c in cases], [os.urandom] * B, return_bytes=False)
    print("batch of %d signatures: %d launches, %d bytes each" % (B, ctx.launches - l0, sizes[0]))
stark.close()
ctx.close()
