#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2h_new_$i.json 2> gpurun_out/r2h_new_$i.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --workload bw_c1 > gpurun_out/r2h_c1.json 2> gpurun_out/r2h_c1.err
python - <<'PY'
import json
for n in ("new_1","new_2","c1"):
    try:
        d=json.load(open(f"gpurun_out/r2h_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e ms", round(d["e2e"].get("ms_per_step",0),3), d["precision_guard"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2h_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bw_bwd4 -s 3 -c 1 -o gpurun_out/r2h_bwd4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2h_ncu_bwd4.log 2>&1; echo "ncu bwd4 rc=$?"
