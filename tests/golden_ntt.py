"""Check a full-size NTT result against the CPU oracle's committed checksums and spot values (tests/golden/bench_digests.json,
made by tools/make_bench_digests.py with oracle/zkoracle.c) - the oracle itself needs ~75 s per 2^26 transform."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_against_golden_ntt(X, log_n):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")))["configs4"][str(log_n)]
    a = np.ascontiguousarray(X).view(np.uint64).reshape(-1, 2)
    k = np.arange(1, len(a) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        got = {"sum_lo": int(a[:, 0].sum(dtype=np.uint64)), "sum_hi": int(a[:, 1].sum(dtype=np.uint64)),
               "wsum_lo": int((a[:, 0] * k).sum(dtype=np.uint64)), "wsum_hi": int((a[:, 1] * k).sum(dtype=np.uint64))}
    assert got == {key: gold[key] for key in got}
    for pos, (lo, hi) in gold["spot"].items():
        assert (int(a[int(pos), 0]), int(a[int(pos), 1])) == (lo, hi), pos
