#!/bin/bash
# finer staged work list as the default (8 items per SM, 8 stages): all GPU tests, the fit-call probe, the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3z_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r3z_pytest.log | tail -8
timeout 300 python scripts/e2e_probe.py 2>&1 | grep "^create" | tail -2
C4=1 PINP=1 timeout 300 python scripts/e2e_probe.py 2>&1 | grep "^create" | tail -2
python bench.py > gpurun_out/bench_r3z.json 2> gpurun_out/bench_r3z.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r3z.json").read().strip().split("\n")[-1])
print("value ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "c4", d["config4"]["ms_per_step"], "c4 e2e", d["config4"]["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
PY
