#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_properties.py -m gpu -q -x > gpurun_out/pytest_gpu_prop.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_prop.log | tail -30
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r2r_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bw_bwdL -s 3 -c 1 -o gpurun_out/r2r_bwdL python bench.py --steps 2 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
