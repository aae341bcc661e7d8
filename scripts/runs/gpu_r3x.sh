#!/bin/bash
# lanes-per-state kernels after the three late changes: times by state count, and one ncu capture of forward + backward at dense N = 16
mkdir -p gpurun_out
timeout 300 python scripts/oddn_probe.py 2>&1 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bw_(fwd|bwd)G" -s 4 -c 2 -o gpurun_out/r3x_generic python scripts/dense_probe.py > gpurun_out/r3x_ncu.log 2>&1; echo "ncu rc=$?"
