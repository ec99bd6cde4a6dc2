#!/bin/bash
# copies the round's evidence from gpurun_out/r02e (scratch) into profiles/ (tracked); run in the build container
S=gpurun_out/r02e; D=profiles
cp $S/bench_default.json $D/r02_bench_1gpu.json
for n in 2 4 8; do [ -f $S/bench_${n}gpu.json ] && tail -1 $S/bench_${n}gpu.json > $D/r02_bench_${n}gpu.json; done
for w in C1 C3 C4 C5; do cp $S/bench_$w.json $D/r02_bench_$w.json; done
cp $S/bench_reference.json $D/r02_bench_reference_arm.json
cp $S/launches.csv $D/r02_launches_bench_steps2_warmup3.csv
python tools/n4_counters.py $S/r02_n4_1776.ncu-rep $S/n4_runinfo.json $D/n4_counters.json > /dev/null
python tools/ncu_summary.py $S/launches.csv $S/r02_n4_1776.ncu-rep > $D/r02_ncu_summary_n4.txt 2>&1
ncu -i $S/r02_n4_1776.ncu-rep --page raw --csv > $D/r02_ncu_n4_raw.csv 2>/dev/null
python tools/ncu_kernels.py $S/r02_small.ncu-rep 888 > $D/r02_ncu_kernels_small.md 2>&1
ncu -i $S/r02_small.ncu-rep --page raw --csv > $D/r02_ncu_small_raw.csv 2>/dev/null
python tools/sass_histogram.py > $D/r02_sass_opcodes.md
cat gpurun_out/r02/diag_phases.log gpurun_out/r02/diag_phases8.log > $D/r02_n4_phase_times.txt 2>/dev/null
cat gpurun_out/r02/variants*.log > $D/r02_n4_variant_timings.txt 2>/dev/null
ls -la $D | tail -30
