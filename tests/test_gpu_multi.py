"""Multi-GPU paths on the GPU box.  The four-step NTT's CUDA steps are exercised on ONE GPU by
running every rank's local steps in turn with the all-to-all done as tensor slicing (the guide
forbids emulating ranks as concurrent processes on one GPU); with >= 2 GPUs the real NCCL path
runs as one process per GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zk_stark_tutor_b200 as zk                                   # noqa: E402
from zk_stark_tutor_b200 import ntt_4step as fs                    # noqa: E402
from oracle import cbind as C, field as F                          # noqa: E402
from golden_ntt import check_against_golden_ntt as _check_against_golden_ntt   # noqa: E402

pytestmark = pytest.mark.gpu


def cuda(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("world,log_n", [(2, 6), (8, 12), (8, 16), (4, 20), (8, 22)])
def test_four_step_cuda_steps_emulated_on_one_gpu(world, log_n):
    import torch
    ctx = zk.Context(0)
    eng = fs.CudaEngine(ctx)
    n = 1 << log_n
    L = n // world
    x = C.synth(0x5EED0005, n)
    w = F.primitive_nth_root(n)
    ys = []
    for r in range(world):                                           # steps 1-2 of every rank
        y = eng.ntt(fs.field_pow(w, world), cuda(fs.scatter_cyclic(x, r, world)))
        ys.append(eng.scale(y, fs.field_pow(w, r)))
    ctx.sync()
    blk = L // world
    pieces = []
    for q in range(world):                                           # step 3 as slicing, step 4 on the GPU
        recv = torch.cat([ys[r][q * blk:(q + 1) * blk] for r in range(world)]).contiguous()
        z = eng.ntt_strided(fs.field_pow(w, L), recv, world, blk, blk)
        ctx.sync()
        pieces.append(host(z).reshape(world, blk, 2))
    assert np.array_equal(fs.gather_natural(pieces), C.ntt(w, x))
    ctx.close()


@pytest.mark.parametrize("world,log_n,inverse", [(2, 6, False), (8, 12, False), (2, 14, False), (8, 16, True), (4, 20, False), (8, 22, False),
                                                  (16, 20, False), (8, 26, False)])
def test_four_step_fused_exchange_emulated_on_one_gpu(world, log_n, inverse):
    """zkb_ntt4_*: every rank of the fused four-step transform (twiddle + exchange stores inside the last local pass, then the
    register-radix cross stage) in ONE process on one GPU - the same kernels and the same peer-pointer stores the multi-GPU run
    uses, with the barrier reduced to stream order.  2^26 / 8 ranks is BASELINE configs[4] at full size."""
    import torch
    ctx = zk.Context(0)
    n = 1 << log_n
    L = n // world
    x = C.synth(0x5EED0005, n)
    w = F.primitive_nth_root(n)
    plans = [fs.Ntt4Plan(ctx, r, world, L) for r in range(world)]
    fs.connect_local(plans)
    xs = [cuda(fs.scatter_cyclic(x, r, world)) for r in range(world)]
    outs = [torch.empty_like(t) for t in xs]
    for rep in range(2):                                              # twice: both halves of the double-buffered receive side
        fs.run_local(plans, w, xs, outs, inverse)
    ctx.sync()
    got = fs.gather_natural([host(o).reshape(world, L // world, 2) for o in outs])
    del xs, outs
    if log_n == 26 and not inverse:
        # full size: against the CPU oracle's committed checksums and spot values of this very transform (tests/golden/bench_digests.json,
        # made by tools/make_bench_digests.py with oracle/zkoracle.c; the oracle needs ~75 s per 2^26 transform)
        _check_against_golden_ntt(got, log_n)
    else:
        assert np.array_equal(got, C.ntt(w, x, inverse=inverse))
    for p in plans:
        p.close()
    ctx.close()


def test_lde_commit_batch_abi():
    """zkb_lde_commit_batch (configs[3] through the C ABI): columns dealt over the contexts of one process; roots == the oracle's"""
    import ctypes
    import torch
    from oracle import proof_stream as PS, fastfri
    from oracle.fri import FRI as OFRI
    ngpu = max(1, min(torch.cuda.device_count(), 4))
    ctxs = [zk.Context(d, stream="own") for d in range(ngpu)] + [zk.Context(0, stream="own")]     # two contexts on GPU 0 as well
    log_n, ncols = 13, 7
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    fri = zk.FRI(F.GENERATOR, w, n, 4, 64, ctxs[0])
    cols = [np.ascontiguousarray(C.synth(0x5EED0004 + c, n // 4)) for c in range(ncols)]
    rounds = fri.num_rounds()
    roots = np.zeros((ncols, rounds, 64), dtype=np.uint8)
    ca = (ctypes.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    pa = (ctypes.c_void_p * ncols)(*[c.ctypes.data for c in cols])
    ctxs[0].check(ctxs[0].lib.zkb_lde_commit_batch(ca, len(ctxs), ctypes.byref(fri.params), pa, n // 4, ncols,
                                                   roots.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
    for c in range(ncols):
        ofri = OFRI(F.GENERATOR, w, n, 4, 64)
        _, trees, _ = fastfri.commit(ofri, C.coset_lde(w, n, F.GENERATOR, cols[c]), PS.IndependentProofStream())
        assert [bytes(roots[c, r]) for r in range(rounds)] == [t.root for t in trees]
    for c in ctxs:
        c.close()


def test_ntt_strided_ragged_counts():
    import torch
    ctx = zk.Context(0)
    eng = fs.CudaEngine(ctx)
    for n, count, stride in ((8, 5, 7), (4, 1000, 1000), (2, 3000, 3001), (64, 130, 130)):
        w = F.primitive_nth_root(n)
        x = C.synth(9, n * stride)
        z = host(eng.ntt_strided(w, cuda(x), n, stride, count))     # ctx runs on torch's stream: ordered
        for q in (0, 1, count // 2, count - 1):
            assert np.array_equal(z[q::stride][:n], C.ntt(w, np.ascontiguousarray(x[q::stride][:n])))
    ctx.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, log_n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = zk.Context(rank)
    n = 1 << log_n
    x = C.synth(0x5EED0005, n)
    w = F.primitive_nth_root(n)
    xl = torch.from_numpy(fs.scatter_cyclic(x, rank, world).view(np.int64)).cuda()
    piece = fs.ntt_4step(fs.CudaEngine(ctx), w, xl, rank, world)
    ctx.sync()
    plan = fs.Ntt4Plan(ctx, rank, world, n // world)                 # the native path: exchange fused into the last local pass, CUDA IPC
    plan.connect_group()
    for _ in range(3):
        fused = fs.ntt_4step_fused(plan, w, xl)
    torch.cuda.synchronize()
    assert torch.equal(fused.reshape(-1, 2), piece.reshape(-1, 2)), "fused exchange != NCCL all-to-all path"
    q.put((rank, piece.cpu().numpy().view(np.uint64).tobytes()))
    dist.barrier()
    plan.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_four_step_nccl_all_gpus():
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (the single-GPU emulation above covers the CUDA steps)")
    log_n = 26 if world >= 8 else 22
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n = 1 << log_n
    L = n // world
    pieces = [np.frombuffer(res[r], dtype=np.uint64).reshape(world, L // world, 2) for r in range(world)]
    got = fs.gather_natural(pieces)
    if log_n == 26:
        _check_against_golden_ntt(got, log_n)          # the CPU oracle's committed checksums + spot values of this transform
    else:
        assert np.array_equal(got, C.ntt(F.primitive_nth_root(n), C.synth(0x5EED0005, n)))
