#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err; echo "bench rc=$?"
tail -12 gpurun_out/r2e_bench1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench1.json"))
for k in ("value","ms_per_step","gpu_launches","clocks","peaks","e2e","precision_guard"):
    print(k, d.get(k))
print("roofline", {k:v for k,v in d["roofline"].items() if k!="phases"})
print("phases", d["roofline"]["phases"])
for k in ("h2d_probe","config4","strong","parity","score_sharded","vq_encode_sharded","vq_encode","lbg","score","score_n16","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "backward_mass" -s 2>&1 | grep -E "hand-overs|passed|failed"
