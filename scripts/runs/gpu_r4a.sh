#!/bin/bash
# k_repack_blocks4 with one MATCH per step instead of eleven ballots and 8-byte codeword fetches: A/B of the fit call
# (pipelined and plain create), the scorer, then all GPU tests
mkdir -p gpurun_out
for v in "" rbase ""; do
  echo "== variant '${v}'"
  if [ -n "$v" ]; then export HMMB_LIB_PATH=$PWD/hmm_training_b200/libhmmb200_$v.so; else unset HMMB_LIB_PATH; fi
  timeout 200 python scripts/e2e_probe.py 2>&1 | tail -3 | cut -c1-200
  PIPE=0 timeout 200 python scripts/e2e_probe.py 2>&1 | tail -2 | cut -c1-200
  timeout 200 python scripts/score_stage_probe.py 2>&1 | tail -2
done
unset HMMB_LIB_PATH
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r4a_pytest.log | tail -5
