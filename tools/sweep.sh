#!/bin/bash
# The metric's whole size range on one GPU: LDE+FRI-commit at 2^16..2^24 and the standalone NTT (forward + inverse) at 2^16..2^26.
# usage (on the GPU box): bash tools/sweep.sh > gpurun_out/r02_sweep.jsonl
for l in 16 18 20 22 24; do python bench.py --log-n $l --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null; done
for l in 16 18 20 22 24 25 26; do python bench.py --workload ntt --log-n $l --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null; done
