import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zk_stark_tutor_b200 as zk
from zk_stark_tutor_b200 import ntt_4step as fs
from oracle import cbind as C, field as F
ctx = zk.Context(0); eng = fs.CudaEngine(ctx)
for n, count, stride in ((64,64,64),(64,128,128),(64,130,130),(64,2,2),(64,2,130),(32,130,130),(128,40,40),(16,300,300)):
    w = F.primitive_nth_root(n)
    x = C.synth(9, n * stride)
    z = eng.ntt_strided(w, torch.from_numpy(x.view(np.int64)).cuda(), n, stride, count); ctx.sync()
    z = z.cpu().numpy().view(np.uint64)
    bad = [q for q in range(count) if not np.array_equal(z[q::stride][:n], C.ntt(w, np.ascontiguousarray(x[q::stride][:n])))]
    print(n, count, stride, "bad:", bad[:10], len(bad))
