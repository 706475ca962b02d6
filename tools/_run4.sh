python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest_gpu.log 2>&1; tail -3 gpurun_out/s4_pytest_gpu.log
for l in 16 20 22 24 26; do python bench.py --workload ntt --log-n $l --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/s4_ntt_$l.json; done
ZKB_NTT_NO_TWX=1 python bench.py --workload ntt --log-n 24 --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/s4_ntt_24_notwx.json
python bench.py --no-cpu-baseline --no-sub-records 2>/dev/null > gpurun_out/s4_bench.json
python bench.py --workload signatures --no-cpu-baseline 2>/dev/null > gpurun_out/s4_sig.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s4_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, round(d["ms_per_step"],4), round(d["value"],1), {k:round(v.get("ms_per_step",0),3) for k,v in d.get("kernels",{}).items()})
    except Exception as e: print(f,"ERR",e)
PY
