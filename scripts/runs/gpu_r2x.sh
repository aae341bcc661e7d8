#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_properties.py -m gpu -q -x > gpurun_out/pytest_gpu_prop.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_prop.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq_assign -s 3 -c 1 -o gpurun_out/r2x_vq python scripts/vq_probe.py > gpurun_out/r2x_ncu_vq.log 2>&1; echo "ncu vq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq_assign -s 700 -c 1 -o gpurun_out/r2x_lbg python scripts/lbg_probe.py > gpurun_out/r2x_ncu_lbg.log 2>&1; echo "ncu lbg rc=$?"
