#!/bin/bash
# k_bw_bwdG: the two group reductions of a step in one interleaved shuffle pass — A/B on dense models, then all GPU tests
mkdir -p gpurun_out
for v in "" gbase ""; do
  echo "== variant '${v}'"
  if [ -n "$v" ]; then export HMMB_LIB_PATH=$PWD/hmm_training_b200/libhmmb200_$v.so; else unset HMMB_LIB_PATH; fi
  timeout 300 python scripts/dense_probe.py 2>&1 | tail -2
done
unset HMMB_LIB_PATH
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3w_pytest.log
