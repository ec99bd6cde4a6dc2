#!/bin/bash
# Final evidence of the round (the ncu --set full captures of r02_evidence.sh stay valid: kernels unchanged)
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $?"; cut -c1-300 $O/bench_default.json
for wl in C1 C3 C4 C5; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc $?"
done
VB_REF_BUDGET_S=25 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cohort-leg > $O/ncu_launches.log 2>&1; echo "launch list rc $?"
python tools/time_big.py > $O/time_big.log 2>&1; tail -6 $O/time_big.log
