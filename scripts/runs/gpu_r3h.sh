#!/bin/bash
for v in main vq3 vq4 main; do
  case $v in main) L=$PWD/hmm_training_b200/libhmmb200.so;; *) L=$PWD/hmm_training_b200/libhmmb200_$v.so;; esac
  echo "== $v"; HMMB_LIB_PATH=$L timeout 300 python scripts/vq_probe.py 2>&1 | tail -2; HMMB_LIB_PATH=$L timeout 300 python scripts/lbg_probe.py 2>&1 | tail -1
done
