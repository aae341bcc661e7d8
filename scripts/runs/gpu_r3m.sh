#!/bin/bash
# ncu of k_bw_bwdL<16> after the partly-filled-block fix (config 4)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r3m_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bw_bwdL -s 3 -c 1 -o gpurun_out/r3m_bwdL python bench.py --steps 2 --warmup 3 --no-extras --workload bw_c4 > gpurun_out/r3m_ncu.log 2>&1; echo "ncu rc=$?"
