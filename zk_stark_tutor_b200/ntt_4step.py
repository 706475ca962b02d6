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


# ---- the native path: exchange fused into the last local pass (csrc/ntt4.cu, zkb_ntt4_*) -------------------------
class Ntt4Plan:
    """zkb_ntt4: rank `rank` of a `world`-rank four-step transform of world * n_local values on ctx's GPU.  The twiddle and the
    exchange are part of the last pass of the local transform: its stores land in the receiving GPU's HBM over NVLink."""

    def __init__(self, ctx, rank, world, n_local):
        import ctypes
        self.ctx, self.rank, self.world, self.n_local = ctx, rank, world, n_local
        self.h = ctypes.c_void_p()
        ctx.check(ctx.lib.zkb_ntt4_create(ctx.h, rank, world, n_local, ctypes.byref(self.h)))

    def export(self):
        import ctypes
        buf = (ctypes.c_uint8 * 128)()
        self.ctx.check(self.ctx.lib.zkb_ntt4_export(self.h, buf))
        return bytes(buf)

    def connect_ipc(self, all_handles):
        import ctypes
        raw = b"".join(all_handles)
        assert len(raw) == 128 * self.world
        self.ctx.check(self.ctx.lib.zkb_ntt4_connect_ipc(self.h, (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)))

    def connect_group(self, group=None):
        """one process per GPU: all-gather the IPC handles over torch.distributed (any backend), then open the peers' buffers"""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return
        backend = dist.get_backend(group)
        dev = torch.device("cuda", self.ctx.device) if backend == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(self.export()), dtype=torch.uint8).to(dev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=group)
        self.connect_ipc([bytes(p.cpu().numpy().tobytes()) for p in parts])

    def scatter(self, w, x_local, inverse=False):
        from .context import Vec, le16
        v = Vec(x_local)
        assert v.n == self.n_local
        self.ctx.check(self.ctx.lib.zkb_ntt4_scatter(self.h, le16(w), 1 if inverse else 0, v.ptr))
        return v

    def finish(self, out):
        from .context import Vec
        v = Vec(out)
        assert v.n == self.n_local
        self.ctx.check(self.ctx.lib.zkb_ntt4_finish(self.h, v.ptr))
        return out

    def close(self):
        if self.h is not None and self.h.value:
            self.ctx.lib.zkb_ntt4_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def connect_local(plans):
    """every rank in this process (one context per GPU, or several ranks on one GPU)"""
    import ctypes
    arr = (ctypes.c_void_p * len(plans))(*[p.h for p in plans])
    plans[0].ctx.check(plans[0].ctx.lib.zkb_ntt4_connect_local(arr, len(plans)))


def run_local(plans, w, xs, outs, inverse=False):
    """zkb_ntt4_run: scatter on every rank, events, finish on every rank; xs / outs: per-rank CUDA tensors (n_local, 2)"""
    import ctypes
    from .context import le16
    n = len(plans)
    pa = (ctypes.c_void_p * n)(*[p.h for p in plans])
    xa = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
    oa = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    plans[0].ctx.check(plans[0].ctx.lib.zkb_ntt4_run(pa, n, le16(w), 1 if inverse else 0, xa, oa))
    return outs


def ntt_4step_fused(plan, w, x_local, out=None, group=None, inverse=False):
    """One process per GPU (torch.distributed, NCCL): scatter -> a one-element all-reduce on the context's stream (the
    stream-ordered barrier: it completes once every rank's scatter kernels have) -> finish.  No host synchronisation.
    The context must run on torch's current stream.  Returns (world, L/world, 2) like ntt_4step."""
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty_like(x_local)
    plan.scatter(w, x_local, inverse)
    if plan.world > 1:
        token = getattr(plan, "_token", None)
        if token is None:
            token = plan._token = torch.zeros(1, dtype=torch.int32, device=x_local.device)
        dist.all_reduce(token, group=group)
    plan.finish(out)
    return out.view(plan.world, plan.n_local // plan.world, 2)
