#!/bin/bash
mkdir -p gpurun_out
echo "== c4 pinned obs, pageable params"; C4=1 HMMB_TIMING=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -25
echo "== c4 pinned everything"; C4=1 PINP=1 HMMB_TIMING=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -25
