"""Four-step NTT across ranks: the exchange plumbing (cyclic scatter -> local NTT -> twiddle ->
all-to-all -> cross-rank NTT -> gather) on CPU with gloo, world sizes 2 and 4.  The local steps
use an oracle-backed engine HERE ONLY (the CUDA engine is covered on the GPU box by
tests/test_gpu_multi.py); the result must equal the oracle's single NTT bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleEngine:
    def ntt(self, root, x):
        from oracle import cbind as C
        return C.ntt(root, x)

    def scale(self, x, factor):
        from oracle import cbind as C, ntt as N
        return C.to_arr(N.scale(C.from_arr(x), factor))

    def ntt_strided(self, root, x, n, stride, count):
        from oracle import cbind as C
        out = np.empty_like(x)
        for q in range(count):
            out[q::stride][:n] = C.ntt(root, np.ascontiguousarray(x[q::stride][:n]))
        return out

    def all_to_all(self, x, group=None):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.int64))
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=group)
        return out.numpy().view(np.uint64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log_n, q):
    import torch.distributed as dist
    from oracle import cbind as C, field as F
    from zk_stark_tutor_b200 import ntt_4step as fs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    x = C.synth(0x5EED0005, n)
    w = F.primitive_nth_root(n)
    piece = fs.ntt_4step(OracleEngine(), w, fs.scatter_cyclic(x, rank, world), rank, world)
    q.put((rank, np.ascontiguousarray(piece).tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 6), (4, 10), (2, 13)])
def test_four_step_matches_single_ntt(world, log_n):
    from oracle import cbind as C, field as F
    from zk_stark_tutor_b200 import ntt_4step as fs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 1 << log_n
    L = n // world
    pieces = [np.frombuffer(res[r], dtype=np.uint64).reshape(world, L // world, 2) for r in range(world)]
    got = fs.gather_natural(pieces)
    x = C.synth(0x5EED0005, n)
    assert np.array_equal(got, C.ntt(F.primitive_nth_root(n), x))


def test_world_one_is_plain_ntt():
    from oracle import cbind as C, field as F
    from zk_stark_tutor_b200 import ntt_4step as fs
    x = C.synth(3, 64)
    w = F.primitive_nth_root(64)
    out = fs.ntt_4step(OracleEngine(), w, fs.scatter_cyclic(x, 0, 1), 0, 1)
    assert np.array_equal(fs.gather_natural([out]), C.ntt(w, x))
