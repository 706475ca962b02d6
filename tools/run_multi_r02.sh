#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8xB200 box (gpurun --gpus 8 -- bash tools/run_multi_r02.sh):
#   1. the NCCL / CUDA-IPC parity tests (four-step NTT at 2^26 over 8 GPUs, fused exchange == NCCL all-to-all path == oracle)
#   2. bench.py's default line under torchrun at N = 2, 4, 8: configs[2] weak scaling with the configs[3] (64 x 2^22 columns, strong)
#      and configs[4] (one 2^26 NTT over N GPUs) sub-records, each parity-checked against the oracle digests before timing
#   3. the proof-batch workload at N = 8
out=gpurun_out/r02_multi_gpu_8xB200.jsonl
: > $out
nvidia-smi -L > gpurun_out/r02_multi_gpu_devices.txt
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_multi_8gpu.log
port=29600
for n in 2 4 8; do
  port=$((port+1))
  NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r02_multi_err_$n.log >> $out || tail -5 gpurun_out/r02_multi_err_$n.log
done
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 8 --workload proofs --proofs 8192 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_multi_err_proofs.log >> $out || tail -5 gpurun_out/r02_multi_err_proofs.log
python - <<'EOF'
import json
for l in open("gpurun_out/r02_multi_gpu_8xB200.jsonl"):
    d = json.loads(l)
    print(d["n_gpus"], d["metric"][:40], round(d["ms_per_step"], 3), round(d["value"], 1), d["unit"], "e2e", round(d["e2e"].get("ms_per_step", 0), 3))
    for k, v in d.get("configs", {}).items():
        print("   ", k, round(v["ms_per_step"], 3), round(v["value"], 1), "e2e", round(v["e2e"]["ms_per_step"], 3), v["parity"]["checked_before_timing"] is not None)
EOF
