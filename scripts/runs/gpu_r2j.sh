#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
HMMB_LTR_NO_PREFETCH=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r2j_c4_nopre.json 2> gpurun_out/r2j_c4_nopre.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r2j_c4_pre.json 2> gpurun_out/r2j_c4_pre.err
HMMB_LTR_NO_PREFETCH=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_n8 > gpurun_out/r2j_n8_nopre.json 2> gpurun_out/r2j_n8_nopre.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_n8 > gpurun_out/r2j_n8_pre.json 2> gpurun_out/r2j_n8_pre.err
python - <<'PY'
import json
for n in ("c4_nopre","c4_pre","n8_nopre","n8_pre"):
    try:
        d=json.load(open(f"gpurun_out/r2j_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e ms", round(d["e2e"].get("ms_per_step",0),3))
    except Exception as e: print(n, "ERR", e)
PY
