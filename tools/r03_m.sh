#!/bin/bash
# 8-GPU box: configs[2] (1 GPU), configs[3] (1 and 8 GPUs), configs[1] at 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 python bench.py --model m --blocks 10 --loss arc --batch 256 --steps 10 --warmup 3 > gpurun_out/r03m_m_arc_b256_n1.json 2> gpurun_out/r03m_m.err
CUDA_VISIBLE_DEVICES=0 python bench.py --model l --blocks 5 --loss arc --ragged --seconds 8 --batch 64 --steps 10 --warmup 3 > gpurun_out/r03m_l_arc_ragged_b64_n1.json 2> gpurun_out/r03m_l1.err
$TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --model l --blocks 5 --loss arc --ragged --seconds 8 --batch 64 --steps 10 --warmup 3 > gpurun_out/r03m_l_arc_ragged_b64_n8.json 2> gpurun_out/r03m_l8.err
$TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r03m_s_n8.json 2> gpurun_out/r03m_s8.err
$TR --nproc-per-node 2 --master-port 29803 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03m_s_n2.json 2> gpurun_out/r03m_s2.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03m_s_n1.json 2> gpurun_out/r03m_s1.err
for f in gpurun_out/r03m_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -2 gpurun_out/r03m_*.err | tail -20
