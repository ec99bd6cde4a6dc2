#!/bin/bash
# times the whole path on the bench batch for the default library and every tuning build
python tools/run_stage.py all 1776 3 2>&1 | tail -1
for so in vent_analysis_b200/lib/variants/*.so; do
  VENTB200_LIB=$PWD/$so python tools/run_stage.py all 1776 3 2>&1 | tail -1
done
