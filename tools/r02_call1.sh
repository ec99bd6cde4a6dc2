#!/bin/bash
# GPU call: parity tests, tuning-variant timings, per-level phase diagnostics (one box, one GPU)
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu.log
python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee $O/variants.log
for so in vent_analysis_b200/lib/variants/*.so; do
  VENTB200_LIB=$PWD/$so timeout 300 python tools/run_stage.py all 1776 3 2>&1 | tail -1 | tee -a $O/variants.log
done
python tools/diag_phases.py - 3 2>&1 | tee $O/diag_phases.log
python tools/diag_phases.py $PWD/vent_analysis_b200/lib/variants/base.so 3 2>&1 | tee -a $O/diag_phases.log
