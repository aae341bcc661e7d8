#!/bin/bash
# round 2, call B: all GPU tests (incl. the full-size parity tests), VQ / LBG / score probes, bench A/B
mkdir -p gpurun_out
R1=$PWD/scripts/_build/libhmmb200_r1.so
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python scripts/vq_probe.py 2>&1 | tail -2
timeout 300 python scripts/lbg_probe.py 2>&1 | tail -12
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2b_new_$i.json 2> gpurun_out/r2b_new_$i.err
done
python - <<'PY'
import json
for n in ("new_1","new_2"):
    try:
        d=json.load(open(f"gpurun_out/r2b_{n}.json"))
        ph=d["roofline"]["phases"]
        print(n, "ms/iter %.3f"%d["ms_per_step"], {k:round(v["ms_per_launch"],3) for k,v in ph.items()}, "frac %.3f"%d["roofline"]["frac"], "e2e ms", round(d["e2e"].get("ms_per_step",0),2), d["precision_guard"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 300 python scripts/vq_probe.py > gpurun_out/r2b_vq_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq_assign -s 3 -c 1 -o gpurun_out/r2b_vq python scripts/vq_probe.py > gpurun_out/r2b_ncu_vq.log 2>&1; echo "ncu vq rc=$?"
