"""One NTT spread over the g GPUs of a box: the four-step algorithm with ONE all-to-all over
NVLink (NCCL), SURVEY.md 8e.2 / BASELINE configs[4].  One process per GPU.

Index algebra (N = g*L, w a primitive N-th root, ntt.rs:7-49 semantics X[k] = sum_n x[n] w^(nk)):
write n = r + g*m (r = rank, m < L) and k = k2 + L*k1 (k2 < L, k1 < g).  Then

    X[k2 + L*k1] = sum_r  w_g^(r*k1) * [ w^(r*k2) * sum_m x[r + g*m] * w_L^(m*k2) ]

  1. rank r holds the CYCLIC slice x[r + g*m] and runs a local L-point NTT with root w^g
  2. local twiddle by w^(r*k2)                     (Polynomial::scale with factor w^r)
  3. all-to-all: the k2 range is cut into g blocks, block q goes to rank q   <- the only exchange
  4. rank q runs L/g interleaved g-point NTTs (root w^L) across the g received pieces
and ends up holding X[k2 + L*k1] for k2 in its block, laid out [k1][k2 - q*L/g].

A contiguous (block) distribution on both sides would need a second all-to-all; since the
host<->device copies can scatter / gather with any stride for free, `scatter_cyclic` and
`gather_natural` give natural-order vectors at the host boundary.

The local steps are pluggable (`engine`): the CUDA engine below is the product; the gloo/CPU test
plugs in an oracle-backed engine to check the exchange plumbing without a GPU."""
import numpy as np


class CudaEngine:
    """Local steps on this rank's GPU through the C ABI (device-resident torch tensors)."""

    def __init__(self, ctx):
        import zk_stark_tutor_b200 as zk
        self.zk, self.ctx = zk, ctx

    def ntt(self, root, x):
        return self.zk.ntt(root, x, self.ctx)

    def scale(self, x, factor):
        return self.zk.scale(x, factor, self.ctx)

    def ntt_strided(self, root, x, n, stride, count):
        import torch
        from .context import le16
        out = torch.empty_like(x)
        self.ctx.check(self.ctx.lib.zkb_ntt_strided(self.ctx.h, le16(root), 0, x.data_ptr(), n, stride, count, out.data_ptr()))
        return out

    def all_to_all(self, x, group=None):
        import torch
        import torch.distributed as dist
        self.ctx.sync()                       # the library's stream -> NCCL's stream
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x, group=group)
        torch.cuda.current_stream().synchronize()
        return out


def field_pow(w, e):
    from .field import Field
    return Field().pow(w, e)


def ntt_4step(engine, w, x_local, rank, world, group=None):
    """x_local: this rank's cyclic slice x[rank + world*m] ((L, 2) 64-bit array / tensor).
    Returns (world, L/world, 2): X[k2 + L*k1] at [k1][k2 - rank*L/world]."""
    L = x_local.shape[0]
    assert L % world == 0 and (L & (L - 1)) == 0 and (world & (world - 1)) == 0
    if world == 1:
        return engine.ntt(w, x_local).reshape(1, L, 2)
    y = engine.ntt(field_pow(w, world), x_local)             # 1. local L-point NTT, root w^g
    y = engine.scale(y, field_pow(w, rank))                  # 2. y[k2] *= w^(rank*k2)
    r = engine.all_to_all(y, group)                          # 3. piece q of y -> rank q; r[n1][k2'] from rank n1
    z = engine.ntt_strided(field_pow(w, L), r, world, L // world, L // world)   # 4. g-point NTTs across n1
    return z.reshape(world, L // world, 2)


def scatter_cyclic(x, rank, world):
    """The slice of a natural-order vector that `rank` owns (host numpy (N, 2))."""
    return np.ascontiguousarray(x[rank::world])


def gather_natural(pieces):
    """pieces[q] = rank q's result (world, L/world, 2) as numpy -> the natural-order (N, 2) vector."""
    world = len(pieces)
    blk = pieces[0].shape[1]
    L = blk * world
    out = np.empty((world * L, 2), dtype=np.uint64)
    for q, p in enumerate(pieces):
        for k1 in range(world):
            out[k1 * L + q * blk:k1 * L + (q + 1) * blk] = p[k1]
    return out
