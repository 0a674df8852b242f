#!/bin/bash
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | grep -E "^E  |passed|failed|Error|^tests.*(Error|FAILED)" | head -20
for e in 1 0; do TN_PROLOG_TC=$e python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prolog_tc=$e', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])"; done
