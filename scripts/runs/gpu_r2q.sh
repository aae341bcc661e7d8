#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
for v in base new base2 new2; do
  case $v in base|base2) L=$PWD/hmm_training_b200/libhmmb200_base.so;; *) L=$PWD/hmm_training_b200/libhmmb200.so;; esac
  HMMB_LIB_PATH=$L timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r2q_$v.json 2> gpurun_out/r2q_$v.err
done
python - <<'PY'
import json
for n in ("base","new","base2","new2"):
    try:
        d=json.load(open(f"gpurun_out/r2q_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", round(d["e2e"]["ms_per_step"],3), d["precision_guard"])
    except Exception as e: print(n, "ERR", e)
PY
