#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "vq or lbg or code_vector or observations or roundtrip" > gpurun_out/pytest_gpu_vq.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_vq.log | tail -10
timeout 300 python scripts/vq_probe.py 2>&1 | tail -5
timeout 300 python scripts/lbg_probe.py 2>&1 | tail -4
