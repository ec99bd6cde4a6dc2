#!/bin/bash
for p in 50 67 75 84 92; do echo "== VB_PART_SPLIT=$p"; VB_PART_SPLIT=$p python tools/run_stage.py all 1776 4 2>&1 | tail -1; done
