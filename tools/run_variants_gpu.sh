#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
T=${1:-12}
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu$T.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu$T.log
python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee $O/variants$T.log
for so in vent_analysis_b200/lib/variants/*.so; do
  VENTB200_LIB=$PWD/$so timeout 300 python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee -a $O/variants$T.log
done
python tools/diag_phases.py - 3 2>&1 | tee $O/diag_phases$T.log
