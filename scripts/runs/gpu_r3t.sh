#!/bin/bash
# dense A on the lanes-per-state kernels: times, then one ncu capture of forward + backward at N = 16
mkdir -p gpurun_out
timeout 300 python scripts/dense_probe.py 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bw_(fwd|bwd)G" -s 4 -c 2 -o gpurun_out/r3t_generic python scripts/dense_probe.py > gpurun_out/r3t_ncu.log 2>&1; echo "ncu rc=$?"
