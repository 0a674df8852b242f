#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python tools/gemm_ab.py 2>&1 | tail -4
python tools/ew_time.py 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['us_per_launch'])"
