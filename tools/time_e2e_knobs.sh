#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu14.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu14.log
python tools/run_stage.py all 1776 4 2>&1 | tail -1
run() { echo "== $*"; env "$@" python tools/time_e2e.py 1776 4 2>&1 | grep -E "stats only|\(e2e\)" | tail -2; }
run A=1
run VB_CHUNK_TAIL=2
run VB_CHUNK_TAIL=4
