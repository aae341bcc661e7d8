#!/bin/bash
# left-to-right kernels: work items per SM (config 4: 1000 words x 16 blocks; 4 -> one item per word, 16 -> two)
mkdir -p gpurun_out
for v in 4 16 4 16 32; do
  echo "== HMMB_LTR_ITEMS_PER_SM=${v}"
  HMMB_LTR_ITEMS_PER_SM=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --workload bw_c4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], {k: round(v['ms_per_launch'], 4) for k, v in d['roofline']['phases'].items()})"
done
