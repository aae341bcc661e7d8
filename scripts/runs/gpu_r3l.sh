#!/bin/bash
# bwdL: alpha-hat rows through cp.async staging, lean arithmetic ahead of the branches, tfull from the forward pass: A/B
mkdir -p gpurun_out
for v in "" noastage noserp ""; do
  echo "== variant '${v}'"
  if [ -n "$v" ]; then export HMMB_LIB_PATH=$PWD/hmm_training_b200/libhmmb200_$v.so; else unset HMMB_LIB_PATH; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --workload bw_c4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], {k: round(v['ms_per_launch'], 4) for k, v in d['roofline']['phases'].items()})"
done
unset HMMB_LIB_PATH
timeout 1200 python -m pytest tests -m gpu -q -x -k "ltr or left or thin or properties or dropin or oracle" > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3l_pytest.log
