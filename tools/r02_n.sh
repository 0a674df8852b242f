#!/bin/bash
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | grep -E "^E  |passed|failed|Error|^tests.*(Error|FAILED)" | head -20
for e in 12 8; do TN_TC_EW7=$e python bench.py --model m --blocks 10 --loss arc --batch 256 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('M ew7=$e', d['ms_per_step'], d['value'])"; done
