#!/bin/bash
# Round-2b multi-GPU records on one box: bash tools/run_multi_r02b.sh "2" (or "2 4 8") [tests] [proofs]
#   the NCCL / CUDA-IPC parity tests (if "tests"), bench.py's default line under torchrun at every N given (configs[2] weak scaling with
#   the configs[3] / configs[4] sub-records, parity-checked before timing), the proof-batch workload at the largest N (if "proofs")
ns=${1:-2}
out=gpurun_out/r02b_multi_gpu.jsonl
: > $out
nvidia-smi -L > gpurun_out/r02b_multi_gpu_devices.txt
if [[ " $* " == *" tests "* ]]; then python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4 | tee gpurun_out/r02b_pytest_multi.log; fi
port=29700
last=1
for n in $ns; do
  port=$((port+1)); last=$n
  NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r02b_multi_err_$n.log >> $out || tail -5 gpurun_out/r02b_multi_err_$n.log
done
if [[ " $* " == *" proofs "* ]]; then
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $last --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $last --workload proofs --proofs 8192 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02b_multi_err_proofs.log >> $out || tail -5 gpurun_out/r02b_multi_err_proofs.log
fi
python - <<'PY'
import json
for l in open("gpurun_out/r02b_multi_gpu.jsonl"):
    d = json.loads(l)
    s = d.get("single_in_flight", {})
    print(d["n_gpus"], d["metric"][:40], round(d["ms_per_step"], 3), round(d["value"], 1), d["unit"], "e2e", round(d["e2e"].get("ms_per_step", 0), 3), "single", round(s.get("ms_per_step", 0), 3))
    for k, v in d.get("configs", {}).items():
        print("   ", k, round(v["ms_per_step"], 3), round(v["value"], 1), "e2e", round(v["e2e"]["ms_per_step"], 3), v["parity"]["checked_before_timing"] is not None)
PY
