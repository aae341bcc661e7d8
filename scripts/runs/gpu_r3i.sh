#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r3i_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3i_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r3i_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bw_(fwd4|bwd4)" -s 6 -c 2 -o gpurun_out/r3i_fwd4_bwd4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r3i_ncu_full.log 2>&1; echo "ncu full rc=$?"
