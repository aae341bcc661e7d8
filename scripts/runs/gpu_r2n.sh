#!/bin/bash
mkdir -p gpurun_out
for v in new w20 abl1 abl2 new2; do
  case $v in new|new2) L=$PWD/hmm_training_b200/libhmmb200.so;; *) L=$PWD/hmm_training_b200/libhmmb200_$v.so;; esac
  HMMB_LIB_PATH=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2n_$v.json 2> gpurun_out/r2n_$v.err
done
python - <<'PY'
import json
for n in ("new","w20","abl1","abl2","new2"):
    try:
        d=json.load(open(f"gpurun_out/r2n_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bw_bwd4 -s 3 -c 1 -o gpurun_out/r2n_bwd4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2n_ncu_bwd4.log 2>&1; echo "ncu bwd4 rc=$?"
