#!/bin/bash
# parity tests of the default build (and of every tuning build under lib/variants), then their stage times
O=gpurun_out/r02; mkdir -p $O; T=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu$T.log 2>&1; echo "pytest rc $?"; tail -1 $O/pytest_gpu$T.log
for so in vent_analysis_b200/lib/variants/*.so; do
  [ -f "$so" ] || continue
  VENTB200_LIB=$PWD/$so timeout 900 python -m pytest tests -m gpu -x -q -k "vdp or anchor or p99 or stress or cohort" > $O/pytest_gpu${T}_$(basename $so .so).log 2>&1; echo "pytest $(basename $so) rc $?"; tail -1 $O/pytest_gpu${T}_$(basename $so .so).log
done
VB_PARTS=1 python tools/run_stage.py vdp 1776 5 2>&1 | tail -1 | tee $O/variants$T.log
for so in vent_analysis_b200/lib/variants/*.so; do
  [ -f "$so" ] || continue
  VB_PARTS=1 VENTB200_LIB=$PWD/$so timeout 300 python tools/run_stage.py vdp 1776 5 2>&1 | tail -1 | tee -a $O/variants$T.log
done
python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee -a $O/variants$T.log
