#!/bin/bash
# k_repack_blocks4: eight ballots for alphabets up to 256 and 8-byte codeword fetches — A/B of the plain create's prepare phase and
# the fit call, one ncu capture of the repack kernel
mkdir -p gpurun_out
for v in "" rbase ""; do
  echo "== variant '${v}'"
  if [ -n "$v" ]; then export HMMB_LIB_PATH=$PWD/hmm_training_b200/libhmmb200_$v.so; else unset HMMB_LIB_PATH; fi
  timeout 200 python scripts/e2e_probe.py 2>&1 | tail -2 | cut -c1-120
  PIPE=0 timeout 200 python scripts/e2e_probe.py 2>&1 | tail -1 | cut -c1-60
done
unset HMMB_LIB_PATH
PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_repack_blocks4 -s 8 -c 1 -o gpurun_out/r4b_repack python scripts/e2e_probe.py > gpurun_out/r4b_ncu.log 2>&1; echo "ncu rc=$?"
