#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02b_tests.log 2>&1
tail -15 gpurun_out/r02b_tests.log
N=6 python tools/parity_repeat.py > gpurun_out/r02b_repeat.log 2>&1
TN_TC_NACC=1 N=4 python tools/parity_repeat.py >> gpurun_out/r02b_repeat.log 2>&1
cat gpurun_out/r02b_repeat.log | tail -8
python __graft_entry__.py --smoke 2>&1 | tail -2
