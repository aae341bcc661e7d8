#!/bin/bash
# final multi-GPU lines of the round on one 8-GPU box: the 2-rank tests, then the full bench at 8 and at 2 ranks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r3r_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/r3r_multi_tests.log
for N in 8 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r3r_bench$N.json 2> gpurun_out/r3r_bench$N.err; echo "bench $N rc=$?"
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/r3r_bench{N}.json").read().strip().split("\n")[-1])
for k in ("value","ms_per_step","gpu_launches","e2e","allreduce"):
    print(k, d.get(k))
print("phases", {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["phases"].items()}, "frac", d["roofline"]["frac"])
for k in ("config4","strong","parity"):
    print(k, json.dumps(d.get(k))[:700])
PY
done
