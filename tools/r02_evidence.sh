#!/bin/bash
# Evidence of the round: parity tests, bench lines of every workload, launch list, ncu captures.
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc $?"; cut -c1-400 $O/bench_default.json
for wl in C1 C3 C4 C5; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc $?"; cut -c1-300 $O/bench_$wl.json
done
VB_REF_BUDGET_S=25 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc $?"
# launch list of the plain bench command (after it exited 0 without ncu above)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cohort-leg > $O/ncu_launches.log 2>&1; echo "launch list rc $?"
# one N4 launch over the bench batch under ncu --set full
VB_PARTS=1 VB_DUMP_INFO=$O/n4_runinfo.json timeout 900 ncu --set full --import-source on --clock-control none -k regex:n4_subject -c 1 -o $O/r02_n4_1776 -f python tools/run_stage.py n4 1776 1 > $O/ncu_n4.log 2>&1; echo "ncu n4 rc $?"
# the small kernels: one launch each
timeout 900 ncu --set full --clock-control none -k regex:"snr_kernel|vdp_stats|vdp_classify4|ci_pack|ci_grow|ci_percentile" -c 6 --launch-skip 12 -o $O/r02_small -f python tools/run_stage.py all 1776 2 > $O/ncu_small.log 2>&1; echo "ncu small rc $?"
ls -la $O
