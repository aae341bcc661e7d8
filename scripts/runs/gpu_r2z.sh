#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu.log | tail -10
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2z_bench1.json 2> gpurun_out/r2z_bench1.err ) 2>&1 | grep real
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z_bench1.json"))
for k in ("value","ms_per_step","gpu_launches","clocks","e2e"):
    print(k, d.get(k))
print("roofline frac", d["roofline"]["frac"], "estep", d["roofline"]["estep"])
print("phases", {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["phases"].items()})
for k in ("config4","strong","vq_encode","lbg","score","lbg_sharded","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
r=json.load(open("gpurun_out/r2z_ref.json")); print("ref", r.get("value"), r.get("cpu_baseline"))
PY
