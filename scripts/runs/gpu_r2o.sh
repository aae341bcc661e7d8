#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
for v in new nopre new2 nopre2; do
  case $v in new|new2) L=$PWD/hmm_training_b200/libhmmb200.so;; nopre2) L=$PWD/hmm_training_b200/libhmmb200_nopre.so;; *) L=$PWD/hmm_training_b200/libhmmb200_$v.so;; esac
  HMMB_LIB_PATH=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2o_$v.json 2> gpurun_out/r2o_$v.err
done
python - <<'PY'
import json
for n in ("new","nopre","new2","nopre2"):
    try:
        d=json.load(open(f"gpurun_out/r2o_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(n, "ERR", e)
PY
echo "== c4 pinned everything"; C4=1 PINP=1 HMMB_TIMING=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12
echo "== c4 pinned obs only"; C4=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -7
