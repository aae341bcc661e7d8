#!/bin/bash
# N = 4 resident work list with exactly sm_count x items_per_sm items (HMMB_BW4_EXACT_ITEMS=0: items rounded up to whole rounds as before)
mkdir -p gpurun_out
for v in 1 0 1; do
  echo "== HMMB_BW4_EXACT_ITEMS=$v"
  HMMB_BW4_EXACT_ITEMS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], {k: round(v['ms_per_launch'], 4) for k, v in d['roofline']['phases'].items()})"
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r4d_pytest.log | tail -5
