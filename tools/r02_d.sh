#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_checkpoint.py tests/test_gpu_model.py -m gpu -q --timeout 900 2>&1 | tail -5
(python tools/gemm_ab.py; TN_TC_FWD_CORR=1 python tools/gemm_ab.py; python tools/gemm_ab.py 19264 256 1536; python tools/gemm_ab.py 19264 1536 128;python tools/gemm_ab.py 77056 512 512; TN_TC_FWD_CORR=1 python tools/gemm_ab.py 77056 512 512) > gpurun_out/r02d_gemm_ab.log 2>&1
cat gpurun_out/r02d_gemm_ab.log
python tools/trace_gemm.py 19264 256 256 3 > gpurun_out/r02d_trace.log 2>&1; tail -30 gpurun_out/r02d_trace.log
