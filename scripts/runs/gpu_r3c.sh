#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thin_states.py -m gpu -q > gpurun_out/pytest_gpu_thin.log 2>&1; echo "pytest thin rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_thin.log | tail -20
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu.log | tail -10
for i in 1 2; do
HMMB_HYP_RANDOM=1 HMMB_HYP_EXAMPLES=150 timeout 900 python -m pytest tests/test_properties.py -m gpu -q > gpurun_out/pytest_gpu_prop_rand$i.log 2>&1; echo "pytest (random $i) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  |p=\(" gpurun_out/pytest_gpu_prop_rand$i.log | tail -12
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r3c_c3.json 2> gpurun_out/r3c_c3.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r3c_c4.json 2> gpurun_out/r3c_c4.err
python - <<'PY'
import json
for n in ("c3","c4"):
    try:
        d=json.load(open(f"gpurun_out/r3c_{n}.json")); ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", round(d["e2e"]["ms_per_step"],3), d["precision_guard"])
    except Exception as e: print(n,"ERR",e)
PY
