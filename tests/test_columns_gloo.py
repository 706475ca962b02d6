"""Host-side multi-rank logic (column partition + root gather) on CPU with the gloo backend,
world_size 2 and 3.  The per-column CUDA work is replaced by a deterministic stand-in (the
oracle's Merkle root of a tiny column) - this test covers only the sharding / collective
plumbing that bench.py and a multi-GPU prover use."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_roots(col, rounds):
    from oracle import merkle as M
    return [M.commit([col * 1000 + r, 7, 8, 9]) for r in range(rounds)]


def _worker(rank, world, port, n_cols, rounds, q):
    import torch.distributed as dist
    from zk_stark_tutor_b200 import columns
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = columns.partition(n_cols, world, rank)
    local = [_fake_roots(c, rounds) for c in mine]
    allr = columns.gather_roots(local, n_cols, world, rank, rounds)
    q.put((rank, mine, allr.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cols", [(2, 8), (2, 5), (3, 7)])
def test_partition_and_gather(world, n_cols):
    from zk_stark_tutor_b200 import columns
    rounds = 3
    # partition covers every column exactly once and is balanced
    allc = sorted(c for r in range(world) for c in columns.partition(n_cols, world, r))
    assert allc == list(range(n_cols))
    sizes = [len(columns.partition(n_cols, world, r)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    assert all(columns.owner(c, world) == r for r in range(world) for c in columns.partition(n_cols, world, r))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, rounds, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros((n_cols, rounds, 64), dtype=np.uint8)
    for c in range(n_cols):
        for r, root in enumerate(_fake_roots(c, rounds)):
            want[c, r] = np.frombuffer(root, dtype=np.uint8)
    for rank, mine, blob in results:
        assert blob == want.tobytes(), "rank %d gathered wrong roots" % rank


def test_single_rank_gather_needs_no_group():
    from zk_stark_tutor_b200 import columns
    local = [_fake_roots(c, 2) for c in range(3)]
    out = columns.gather_roots(local, 3, 1, 0, 2)
    assert out.shape == (3, 2, 64) and bytes(out[2, 1]) == local[2][1]
