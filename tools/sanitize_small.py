"""A small pass over the round-2 kernels for compute-sanitizer (memcheck): device Fiat-Shamir commit with the persistent tail, the
batched provers with device-side framing, the fused four-step NTT (ranks emulated in one process), the opt-in fused NTT + leaf pass.
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zk_stark_tutor_b200 as zk
from zk_stark_tutor_b200 import synth, ntt_4step as fs

G = 85408008396924667383611388730472331217
ctx = zk.Context(0)
field = zk.Field()


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


for log_n in (10, 14, 18):                                   # tail only / tail only / one large layer + tail
    n = 1 << log_n
    w = field.primitive_nth_root(n)
    fri = zk.FRI(G, w, n, 4, 16, ctx)
    ps = zk.SignatureProofStream(b"doc")
    fri.lde_commit(cuda(synth.elements(1, n // 4)), ps).close()
    cw = zk.fast_coset_evaluate(w, n, G, cuda(synth.elements(2, n // 4)), ctx)
    fri.prove(cw, zk.IndependentProofStream())
os.environ["ZKB_NTT_LEAF_FUSION"] = "1"
n = 1 << 18
fri = zk.FRI(G, field.primitive_nth_root(n), n, 4, 16, ctx)
fri.lde_commit(cuda(synth.elements(3, n // 4)), zk.IndependentProofStream()).close()
os.environ.pop("ZKB_NTT_LEAF_FUSION")
# batched provers
n, B, ncc = 1 << 12, 3, 16
w = field.primitive_nth_root(n)
fri = zk.FRI(G, w, n, 4, ncc, ctx)
cws = torch.stack([zk.fast_coset_evaluate(w, n, G, cuda(synth.elements(10 + b, n // 4)), ctx) for b in range(B)]).contiguous()
streams = [zk.SignatureProofStream(b"d%d" % b) for b in range(B)]
handles = (ctypes.c_void_p * B)(*[s.h.value for s in streams])
trees = (ctypes.c_void_p * B)()
ctx.check(ctx.lib.zkb_merkle_build_batch(ctx.h, cws.data_ptr(), n, n, B, trees, handles))
top = np.empty((B, ncc), dtype=np.uint64)
ctx.check(ctx.lib.zkb_fri_prove_batch(ctx.h, ctypes.byref(fri.params), cws.data_ptr(), n, n, B, handles, top.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
idx = np.ascontiguousarray(np.random.RandomState(1).randint(0, n, size=(B, 37)).astype(np.uint64))
ctx.check(ctx.lib.zkb_merkle_open_ps_batch(trees, B, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 37, handles))
for i in range(B - 1, -1, -1):
    ctx.lib.zkb_merkle_free(trees[i])
# four-step NTT, 4 ranks in one process
world, L = 4, 1 << 13
w = field.primitive_nth_root(world * L)
plans = [fs.Ntt4Plan(ctx, r, world, L) for r in range(world)]
fs.connect_local(plans)
xs = [cuda(synth.elements(5, L, start=r, step=world)) for r in range(world)]
outs = [torch.empty_like(t) for t in xs]
fs.run_local(plans, w, xs, outs)
ctx.sync()
for p in plans:
    p.close()
ctx.close()
print("sanitize_small: done")
