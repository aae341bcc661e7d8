#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q -x -k "pipelined or left_to_right or config4" > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_p.log | tail -30
echo "== c4 pinned everything, B pieces"; C4=1 PINP=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -5
echo "== c4 pinned everything, no pieces"; HMMB_NO_B_PIECES=1 C4=1 PINP=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -5
echo "== c4 pinned obs only, B pieces"; C4=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -5
echo "== c4 8 stages"; HMMB_PIPE_STAGES=8 C4=1 PINP=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -5
echo "== c4 2 stages"; HMMB_PIPE_STAGES=2 C4=1 PINP=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -5
