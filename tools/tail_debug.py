"""Per-phase clock stamps of the persistent FRI tail kernel (ZKB_TAIL_DEBUG=1): python tools/tail_debug.py [log_n]"""
import os, sys
os.environ["ZKB_TAIL_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zk_stark_tutor_b200 as zk
from zk_stark_tutor_b200 import synth
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << log_n
ctx = zk.Context(0)
w = zk.Field().primitive_nth_root(n)
G = 85408008396924667383611388730472331217
coeffs = torch.from_numpy(synth.elements(1, n // 4).view(np.int64)).cuda()
fri = zk.FRI(G, w, n, 4, 64, ctx)
for it in range(3):
    print("--- run", it, file=sys.stderr)
    fri.lde_commit(coeffs, zk.IndependentProofStream()).close()
