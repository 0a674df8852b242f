#!/bin/bash
# A/B of programmatic dependent launch (TN_PDL=1) on the benchmark step, same box
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default ', d['ms_per_step'], d['value'])"
TN_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TN_PDL=1', d['ms_per_step'], d['value'])"
done
