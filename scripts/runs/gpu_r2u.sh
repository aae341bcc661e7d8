#!/bin/bash
mkdir -p gpurun_out
for items in 2 4; do for st in 4 6 8; do
  echo "== config 3 e2e: items_per_sm=$items stages=$st"
  HMMB_BW4_ITEMS_PER_SM=$items HMMB_PIPE_STAGES=$st timeout 300 python scripts/e2e_probe.py 2>&1 | grep "^create" | tail -2
done; done
echo "== resident iteration with 4 items per SM"
HMMB_BW4_ITEMS_PER_SM=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2u_items4.json 2> gpurun_out/r2u_items4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_items4.json")); ph=d["roofline"]["phases"]
print("items4", "ms/iter %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in ph.items()}, "e2e", round(d["e2e"]["ms_per_step"],3))
PY
