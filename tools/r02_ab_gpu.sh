#!/bin/bash
# A/B of two builds of the same sources on one box: the GPU suite on the current build, then the bench line of the
# current build, of lib/variants/libventb200_verified.so (VENTB200_LIB), and of the current build again
O=gpurun_out/r02h; mkdir -p $O
timeout 100 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cohort-leg"
timeout 60 $B > $O/bench_new_1.json 2> /dev/null; echo "rc=$?"
VENTB200_LIB=$PWD/vent_analysis_b200/lib/variants/libventb200_verified.so timeout 60 $B > $O/bench_verified.json 2> /dev/null; echo "rc=$?"
timeout 60 $B > $O/bench_new_2.json 2> /dev/null; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_new_1", "bench_verified", "bench_new_2"):
    try:
        d = json.loads(open(f"gpurun_out/r02h/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
