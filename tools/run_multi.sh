#!/bin/bash
# Multi-GPU measurement pass (run on an N-GPU box: gpurun --gpus 8 -- 'bash tools/run_multi.sh 8').
# Writes one JSON line per run under gpurun_out/multi/.
N=${1:-8}
mkdir -p gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() {  # run <name> <nproc> <bench args...>
    local name=$1 np=$2; shift 2
    port=$((port + 1))
    if [ "$np" = 1 ]; then
        timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/multi/$name.json 2> gpurun_out/multi/$name.err
    else
        timeout 600 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > gpurun_out/multi/$name.json 2> gpurun_out/multi/$name.err
    fi
    echo "$name rc=$?"; tail -c 300 gpurun_out/multi/$name.json | head -c 300; echo
}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi/pytest_multi.log 2>&1; tail -2 gpurun_out/multi/pytest_multi.log
for np in 1 2 4 8; do
    [ $np -le $N ] && run columns_n$np $np --workload columns --columns 64 --log-n 22 --steps 3 --warmup 3 --no-cpu-baseline
done
for np in 4 8; do
    [ $np -le $N ] && run weak_n$np $np --steps 5 --warmup 3 --no-cpu-baseline
done
for np in 4 8; do
    [ $np -le $N ] && run ntt4step_n$np $np --workload ntt4step --log-n 26 --steps 5 --warmup 3 --no-cpu-baseline
done
for np in 1 2 4 8; do
    [ $np -le $N ] && run proofs_n$np $np --workload proofs --proofs 8192 --proof-batch 64 --lanes 4 --steps 3 --warmup 3 --no-cpu-baseline
done
