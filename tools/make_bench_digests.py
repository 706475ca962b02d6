#!/usr/bin/env python
"""Golden digests for bench.py's pre-timing parity checks, computed with the ORACLE (CPU) in the build container:
    python tools/make_bench_digests.py [--only cfg2|cfg3|cfg4] -> tests/golden/bench_digests.json
bench.py's B200 arm compares its GPU results with these before it times anything (it never runs the oracle itself).
  configs[2]  sha256 of the proof stream after LDE + FRI commit (16-byte header, R Root objects, the last Codeword) for the
              bench input (seed 0x5EED0003, 2^(log_n-2) coefficients), log_n in 16..24
  configs[3]  the same digest for columns 0, 31, 63 of the 64 x 2^22 batch (seeds 0x5EED0003 + column)
  configs[4]  position-weighted checksums of the 2^26-point NTT of stream 0x5EED0005 (any distribution of the output over ranks
              can accumulate them: sum_k X_lo[k]*(k+1), sum_k X_hi[k]*(k+1) mod 2^64, plus plain sums), and a few spot values
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import cbind as C, field as F, proof_stream as PS, fastfri  # noqa: E402
from oracle.fri import FRI  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
SEED, SEED_NTT = 0x5EED0003, 0x5EED0005


def commit_digest(log_n, seed):
    n = 1 << log_n
    w = F.primitive_nth_root(n)
    cw = C.coset_lde(w, n, F.GENERATOR, C.synth(seed, n // 4))
    ps = PS.IndependentProofStream()
    fastfri.commit(FRI(F.GENERATOR, w, n, 4, 64), cw, ps)
    return hashlib.sha256(ps.digest()).hexdigest()


def checksums(arr, first_index=0, step=1):
    """(n, 2) uint64 -> the four wrap-around sums over global indices first_index + step*i"""
    a = np.ascontiguousarray(arr).view(np.uint64).reshape(-1, 2)
    k = np.uint64(first_index) + np.arange(len(a), dtype=np.uint64) * np.uint64(step) + np.uint64(1)
    with np.errstate(over="ignore"):
        return {"sum_lo": int(a[:, 0].sum(dtype=np.uint64)), "sum_hi": int(a[:, 1].sum(dtype=np.uint64)),
                "wsum_lo": int((a[:, 0] * k).sum(dtype=np.uint64)), "wsum_hi": int((a[:, 1] * k).sum(dtype=np.uint64))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    try:
        gold = json.load(open(OUT))
    except OSError:
        gold = {}
    gold["_made_by"] = "tools/make_bench_digests.py (oracle/zkoracle.c zo_* kernels + oracle/fastfri.py)"
    if args.only in ("", "cfg2"):
        for log_n in (16, 18, 20, 22, 24):
            t0 = time.time()
            gold.setdefault("configs2", {})[str(log_n)] = commit_digest(log_n, SEED)
            print("configs[2] 2^%d: %.1f s" % (log_n, time.time() - t0), flush=True)
            json.dump(gold, open(OUT, "w"), indent=1)
    if args.only in ("", "cfg3"):
        for col in (0, 31, 63):
            t0 = time.time()
            gold.setdefault("configs3", {})[str(col)] = commit_digest(22, SEED + col)
            print("configs[3] column %d: %.1f s" % (col, time.time() - t0), flush=True)
            json.dump(gold, open(OUT, "w"), indent=1)
    if args.only in ("", "cfg4"):
        for log_n in (20, 26):
            t0 = time.time()
            n = 1 << log_n
            x = C.synth(SEED_NTT, n)
            X = C.ntt(F.primitive_nth_root(n), x)
            rec = checksums(X)
            rec["spot"] = {str(k): [int(X[k, 0]), int(X[k, 1])] for k in (0, 1, n // 3, n // 2 + 7, n - 1)}
            gold.setdefault("configs4", {})[str(log_n)] = rec
            print("configs[4] 2^%d: %.1f s" % (log_n, time.time() - t0), flush=True)
            json.dump(gold, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
