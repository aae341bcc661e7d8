#!/bin/bash
# N = 2: the multi-rank tests and the full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest multi rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|skipped" gpurun_out/pytest_gpu_multi.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v_bench2.json 2> gpurun_out/r2v_bench2.err; echo "bench rc=$?"
grep -v "^\[W\|NCCL\|^$" gpurun_out/r2v_bench2.err | tail -15
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_bench2.json"))
for k in ("value","ms_per_step","gpu_launches","clocks","e2e","allreduce"):
    print(k, d.get(k))
print("phases", {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["phases"].items()})
for k in ("h2d_probe","config4","strong","parity","score_sharded","vq_encode_sharded"):
    print(k, json.dumps(d.get(k))[:1200])
PY
