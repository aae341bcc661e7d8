#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r3e_bench$N.json 2> gpurun_out/r3e_bench$N.err; echo "bench rc=$?"
grep -v "^\[W\|NCCL\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r3e_bench$N.err | tail -12
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.load(open(f"gpurun_out/r3e_bench{N}.json"))
for k in ("value","ms_per_step","gpu_launches","clocks","e2e","allreduce"):
    print(k, d.get(k))
print("phases", {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["phases"].items()}, "frac", d["roofline"]["frac"])
for k in ("h2d_probe","config4","strong","parity","score_sharded","vq_encode_sharded","lbg_sharded"):
    print(k, json.dumps(d.get(k))[:1000])
PY
