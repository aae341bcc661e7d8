#!/bin/bash
# final code on two GPUs: the 2-rank tests and the full bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r4e_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/r4e_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4e_bench2.json 2> gpurun_out/r4e_bench2.err; echo "bench 2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4e_bench2.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["phases"].items()})
print("config4", d["config4"]["ms_per_step"], "strong", d["strong"]["ms_per_step"], "parity", d["parity"]["pass"], {k:v.get("max_rel_err") for k,v in d["parity"]["cases"].items()})
PY
