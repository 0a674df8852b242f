#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02c_tests.log 2>&1
tail -25 gpurun_out/r02c_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
head -c 1500 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
