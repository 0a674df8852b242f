#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | grep -E "^E  |passed|failed|Error|^tests.*(Error|FAILED)" | head -20
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02l.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S b64', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['hbm_peak_gb'])"
tail -2 gpurun_out/r02l.err
python tools/sweep.py --batches 32,64,128,256,512,1024,2048 --out gpurun_out/r02l_sweep_s.jsonl 2>&1 | tail -8
TN_RECOMPUTE_U=1 python tools/sweep.py --batches 64,2048 --out gpurun_out/r02l_sweep_s_recompute.jsonl 2>&1 | tail -3
