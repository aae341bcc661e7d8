#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python scripts/score_probe.py 2>&1 | tail -4
HMMB_SCORE_NO_REPLICAS=1 timeout 300 python scripts/score_probe.py 2>&1 | tail -2
timeout 300 python scripts/score_probe.py > gpurun_out/r2d_score_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score4r -s 2 -c 1 -o gpurun_out/r2d_score python scripts/score_probe.py > gpurun_out/r2d_ncu_score.log 2>&1; echo "ncu score rc=$?"
