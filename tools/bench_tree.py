#!/usr/bin/env python
"""Merkle commit latency per tree size (device-resident leaves), for tuning the latency-mode tree.

usage: python tools/bench_tree.py [log_n ...]     (env knobs: ZKB_TREE_SPLIT_LOG, ZKB_TREE_CHUNKS_LOG,
                                                   ZKB_TREE_LEAF_LOG, ZKB_TREE_NODE_LOG)
Prints one line per size: wall us per zkb_merkle_commit call (launches + root D2H + sync) and the
device time of the kernel classes involved (CUDA events on the launching stream).
"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import zk_stark_tutor_b200 as zk  # noqa: E402
from zk_stark_tutor_b200 import synth  # noqa: E402


def main():
    logs = [int(a) for a in sys.argv[1:]] or [9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 22, 24]
    ctx = zk.Context(0)
    knobs = {k: os.environ.get(k) for k in ("ZKB_TREE_SPLIT_LOG", "ZKB_TREE_CHUNKS_LOG", "ZKB_TREE_LEAF_LOG", "ZKB_TREE_NODE_LOG")}
    print("knobs", knobs)
    for ln in logs:
        vals = torch.from_numpy(synth.elements(7, 1 << ln).view(np.int64)).cuda()
        reps = 20 if ln <= 20 else 5
        for _ in range(3):
            root = zk.MerkleRoot.commit(vals, ctx)
        ctx.profile(True, reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            root = zk.MerkleRoot.commit(vals, ctx)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps * 1e6
        prof = ctx.profile_read()
        ctx.profile(False)
        parts = " ".join("%s=%.1fus/%d" % (k, v[0] * 1e3 / reps, v[1] // reps) for k, v in prof.items())
        print("log_n=%2d wall=%8.1f us  %s  root=%s" % (ln, dt, parts, hashlib.sha256(root).hexdigest()[:12]))


if __name__ == "__main__":
    main()
