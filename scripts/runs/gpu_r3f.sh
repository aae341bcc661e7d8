#!/bin/bash
mkdir -p gpurun_out
for v in main w12 w14 pf8 pf20 main2; do
  case $v in main|main2) L=$PWD/hmm_training_b200/libhmmb200.so;; *) L=$PWD/hmm_training_b200/libhmmb200_$v.so;; esac
  HMMB_LIB_PATH=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r3f_$v.json 2> gpurun_out/r3f_$v.err
done
python - <<'PY'
import json
for n in ("main","w12","w14","pf8","pf20","main2"):
    try:
        d=json.load(open(f"gpurun_out/r3f_{n}.json")); ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(n,"ERR",e)
PY
