#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu4.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu4.log
python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee $O/variants4.log
for so in vent_analysis_b200/lib/variants/*.so; do
  VENTB200_LIB=$PWD/$so timeout 300 python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee -a $O/variants4.log
done
python tools/diag_phases.py - 3 2>&1 | tee $O/diag_phases4.log
# one N4 launch over the bench batch under ncu (set full, source counters)
VB_PARTS=1 VB_DUMP_INFO=$O/n4_runinfo.json timeout 900 ncu --set full --import-source on --clock-control none -k regex:n4_subject -c 1 -o $O/r02_n4_1776 -f python tools/run_stage.py n4 1776 1 > $O/ncu_n4.log 2>&1; echo "ncu rc $?"; tail -2 $O/ncu_n4.log
