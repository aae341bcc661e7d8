#!/bin/bash
# finer work list for the staged first E-step of a pipelined N = 4 create: fit-call time by items per SM and stage count,
# then the pipelined-create tests
mkdir -p gpurun_out
for items in 2 6 8 12 16; do for st in 4 8; do
  echo "== config 3 fit call: stage items per SM = $items (2 = the resident list), stages = $st"
  HMMB_BW4_STAGE_ITEMS_PER_SM=$items HMMB_PIPE_STAGES=$st timeout 300 python scripts/e2e_probe.py 2>&1 | grep "^create" | tail -2
done; done
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r3y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3y_pytest.log
