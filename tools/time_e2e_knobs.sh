#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu11.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest_gpu11.log
run() { echo "== $*"; env "$@" python tools/time_e2e.py 1776 4 2>&1 | grep -E "stats only|\(e2e\)" | tail -2; }
run VB_N4_STREAMS=3
run VB_N4_STREAMS=2
run VB_N4_STREAMS=4
run VB_N4_STREAMS=3 VB_CHUNK_PER=111
run VB_N4_STREAMS=4 VB_CHUNK_PER=111
run VB_N4_STREAMS=4 VB_CHUNK_PER=74
