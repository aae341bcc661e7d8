#!/bin/bash
# fresh random examples for the property tests after the flush change of k_bw_bwdL (derandomised in the suite)
mkdir -p gpurun_out
for i in 1 2; do
HMMB_HYP_RANDOM=1 HMMB_HYP_EXAMPLES=500 timeout 1500 python -m pytest tests/test_properties.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r3q_prop_$i.log 2>&1; echo "sweep $i rc=$?"
tail -3 gpurun_out/r3q_prop_$i.log
done
