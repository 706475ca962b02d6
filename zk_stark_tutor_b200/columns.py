"""Independent trace columns / codewords across the GPUs of one box (BASELINE
configs[3]): every column is an independent LDE -> Merkle -> FRI commit, so columns are dealt
round-robin to ranks (one process per GPU, its own context and stream) and NO field data
crosses NVLink - only the 64-byte roots are gathered (torch.distributed, NCCL on GPUs / gloo in
the CPU tests).  This mirrors how Stark::prove treats its registers (stark.rs:373-381: one
fast_coset_evaluate + MerkleRoot::commit per register, independent of each other)."""
import numpy as np


def partition(n_cols, world, rank):
    """Column indices owned by `rank` (round-robin: balanced for any n_cols)."""
    return list(range(rank, n_cols, world))


def owner(col, world):
    return col % world


def lde_commit_columns(fri, columns, make_stream, ctx=None):
    """LDE + FRI commit of each local coefficient column.  `columns`: iterable of (n_coeffs, 2)
    arrays / CUDA tensors.  Returns [(roots [R x 64 bytes], proof-stream digest)] per column."""
    out = []
    for col in columns:
        ps = make_stream()
        layers = fri.lde_commit(col, ps)
        roots = [layers.root(r) for r in range(len(layers))]
        layers.close()
        out.append((roots, ps.digest()))
    return out


class ColumnPipeline:
    """Several columns in flight on ONE GPU: `lanes` contexts (each with its own CUDA stream)
    driven by `lanes` host threads.  A single column's FRI commit is a serial chain (round r+1
    needs the challenge derived from round r's root) whose small rounds are latency-bound and
    leave most SMs idle; independent columns fill them.  The C calls release the GIL."""

    def __init__(self, device, fri_params, lanes=3):
        import zk_stark_tutor_b200 as zk
        self.zk = zk
        self.ctxs = [zk.Context(device, stream="own") for _ in range(lanes)]
        for c in self.ctxs:      # with several lanes the copy engines overlap a staged H2D with the other lanes' kernels
            c.check(c.lib.zkb_ctx_zero_copy_inputs(c.h, 0))
            # ... and the lanes hide each other's latency: per-round launches (device Fiat-Shamir, no host hop) instead of the
            # 128-SM persistent tail kernel (64 x 2^22 columns, 4 lanes, one B200: 166 ms with the persistent tail, 159 ms with host hops)
            c.check(c.lib.zkb_ctx_tail_threads(c.h, 0))
        offset, omega, n, ef, ncc = fri_params
        self.fris = [zk.FRI(offset, omega, n, ef, ncc, c) for c in self.ctxs]

    def run(self, columns, make_stream, keep_roots=True):
        """columns: list of device tensors / host arrays.  Returns per column (roots, digest)."""
        import threading
        results = [None] * len(columns)
        errors = []

        def worker(lane):
            try:
                for i in range(lane, len(columns), len(self.ctxs)):
                    ps = make_stream()
                    layers = self.fris[lane].lde_commit(columns[i], ps)
                    roots = [layers.root(r) for r in range(len(layers))] if keep_roots else None
                    layers.close()
                    results[i] = (roots, ps.digest())
            except Exception as e:       # noqa: BLE001 - re-raised in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(self.ctxs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for c in self.ctxs:
            c.sync()
        if errors:
            raise errors[0]
        return results

    def close(self):
        for c in self.ctxs:
            c.close()


def gather_roots(local_roots, n_cols, world, rank, rounds, group=None):
    """All ranks receive every column's roots: (n_cols, rounds, 64) uint8.  `local_roots`:
    list (in partition order) of per-column lists of 64-byte roots.  world == 1 needs no
    process group."""
    mine = partition(n_cols, world, rank)
    assert len(local_roots) == len(mine)
    per_rank = (n_cols + world - 1) // world
    buf = np.zeros((per_rank, rounds, 64), dtype=np.uint8)
    for k, roots in enumerate(local_roots):
        assert len(roots) == rounds
        for r, root in enumerate(roots):
            buf[k, r] = np.frombuffer(root, dtype=np.uint8)
    out = np.zeros((n_cols, rounds, 64), dtype=np.uint8)
    if world == 1:
        out[:] = buf[:n_cols]
        return out
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    for rk, p in enumerate(parts):
        p = p.cpu().numpy()
        for k, col in enumerate(partition(n_cols, world, rk)):
            out[col] = p[k]
    return out
