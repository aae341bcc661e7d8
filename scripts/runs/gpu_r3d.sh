#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
HMMB_HYP_RANDOM=1 HMMB_HYP_EXAMPLES=600 timeout 1500 python -m pytest tests/test_properties.py -m gpu -q > gpurun_out/pytest_gpu_prop_big$i.log 2>&1; echo "pytest (random big $i) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  +Assert|p=\(" gpurun_out/pytest_gpu_prop_big$i.log | tail -12
grep -A12 "Failing test case" gpurun_out/pytest_gpu_prop_big$i.log | tr -d '\n' | sed 's/E  */ /g' | cut -c1-400; echo
done
