#!/bin/bash
# same-box A/B of the benchmark step: the tree of the round's first commit (exported to the git-ignored _prev/) against HEAD, alternating.
# _prev/ is made in the build container with:  mkdir _prev && git archive 24e7abe | tar -x -C _prev && (cd _prev && python -m titanet_b200._build)
# (gpurun ships git-ignored files, so both trees and both libraries travel to the GPU box)
mkdir -p gpurun_out
show() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], 'ms/step', d['value'], 'utt/s   e2e', d['e2e']['value'], '  dominant GEMM us/launch', d['roofline']['us_per_launch'])"; }
for i in 1 2; do
  (cd _prev && python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null) | show "round-start tree (24e7abe):"
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | show "HEAD:                      "
done | tee gpurun_out/r03_ab_same_box.log
