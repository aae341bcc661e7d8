#!/bin/bash
# round 2, call A: parity of the new backward step, A/B against the round-1 library, ncu of bwd4 / vq / score
mkdir -p gpurun_out
R1=$PWD/scripts/_build/libhmmb200_r1.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2a_pytest.log
for i in 1 2; do
  HMMB_LIB_PATH=$R1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2a_old_$i.json 2> gpurun_out/r2a_old_$i.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2a_new_$i.json 2> gpurun_out/r2a_new_$i.err
done
python - <<'PY'
import json
for n in ("old_1","new_1","old_2","new_2"):
    try:
        d=json.load(open(f"gpurun_out/r2a_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.3f"%d["ms_per_step"], {k:round(v["ms_per_launch"],3) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e ms", round(d["e2e"].get("ms_per_step",0),2), d["precision_guard"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2a_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bw_bwd4 -s 3 -c 1 -o gpurun_out/r2a_bwd4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2a_ncu_bwd4.log 2>&1; echo "ncu bwd4 rc=$?"
timeout 300 python scripts/vq_probe.py > gpurun_out/r2a_vq_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq_assign -s 3 -c 1 -o gpurun_out/r2a_vq python scripts/vq_probe.py > gpurun_out/r2a_ncu_vq.log 2>&1; echo "ncu vq rc=$?"
cat gpurun_out/r2a_vq_plain.log
timeout 300 python scripts/score_probe.py > gpurun_out/r2a_score_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 2 -c 1 -o gpurun_out/r2a_score python scripts/score_probe.py > gpurun_out/r2a_ncu_score.log 2>&1; echo "ncu score rc=$?"
cat gpurun_out/r2a_score_plain.log
ls -la gpurun_out/*.ncu-rep
