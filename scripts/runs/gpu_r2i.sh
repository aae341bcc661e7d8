#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4 6; do
  HMMB_BW4_ITEMS_PER_SM=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2i_$v.json 2> gpurun_out/r2i_$v.err
done
python - <<'PY'
import json
for n in ("2","3","4","6"):
    try:
        d=json.load(open(f"gpurun_out/r2i_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e ms", round(d["e2e"].get("ms_per_step",0),3))
    except Exception as e: print(n, "ERR", e)
PY
