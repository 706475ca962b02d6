#!/bin/bash
# Round-2 single-GPU records (run on the GPU box): the driver-comparable bench line, the size sweep, the ncu launch list and two
# full ncu captures (NTT passes; leaf hashing).  usage: bash tools/run_records_r02.sh   -> gpurun_out/r02b_*
O=gpurun_out
python bench.py > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err
bash tools/sweep.sh > $O/r02b_sweep.jsonl
python bench.py --workload signatures --no-cpu-baseline > $O/r02b_signatures.json 2>/dev/null
python bench.py --workload proofs --no-cpu-baseline > $O/r02b_proofs.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > $O/ncu_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_rr -c 3 -f -o $O/r02b_prof_ntt \
    python bench.py --workload ntt --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_leaf8 -c 2 -f -o $O/r02b_prof_leaf \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub-records > $O/ncu_leaf.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b_bench_n1.json").read())
print("bench", d["ms_per_step"], d["e2e"].get("ms_per_step"), d["roofline"]["frac"], d["clocks"])
for k, v in d["configs"].items(): print(k, v["ms_per_step"], v["clocks"])
for ln in open("gpurun_out/r02b_sweep.jsonl"):
    try:
        r = json.loads(ln); print(r["config"].get("log_n"), r["metric"][:24], round(r["ms_per_step"], 4))
    except Exception as e: print("ERR", e)
for f in ("r02b_signatures", "r02b_proofs"):
    r = json.loads(open("gpurun_out/%s.json" % f).read()); print(f, r["value"], r["unit"])
PY
