#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_properties.py -m gpu -q -x > gpurun_out/pytest_gpu_prop.log 2>&1; echo "pytest (derandomised) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_prop.log | tail -6
for i in 1 2 3; do
HMMB_HYP_RANDOM=1 HMMB_HYP_EXAMPLES=120 timeout 900 python -m pytest tests/test_properties.py -m gpu -q > gpurun_out/pytest_gpu_prop_rand$i.log 2>&1; echo "pytest (random $i) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  |p=\(" gpurun_out/pytest_gpu_prop_rand$i.log | tail -12
done
