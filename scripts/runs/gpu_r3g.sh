#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thin_states.py tests/test_properties.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu_thin.log 2>&1; echo "pytest thin+prop rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  +Assert" gpurun_out/pytest_gpu_thin.log | tail -10
bash scripts/runs/gpu_r3f.sh
