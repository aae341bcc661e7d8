#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "generic or seeded or size_limits or denormal or golden or properties or virtual or left_to_right" > gpurun_out/pytest_gpu_gen.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_gen.log | tail -10
# dense A at config 4's shape (10 %): generic kernels
HMMB_FORCE_DENSE_A=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --workload bw_c4 --scale 0.1 > gpurun_out/r3a_dense.json 2> gpurun_out/r3a_dense.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3a_dense.json")); ph=d["roofline"]["phases"]
print("dense c4 x0.1", "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()})
PY
