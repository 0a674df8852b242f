#!/bin/bash
# round-2 first probe: state of the round-1 tree on a fresh box
mkdir -p gpurun_out
python tools/measure_tf32_peak.py > gpurun_out/r02a_tf32_peak.json 2> gpurun_out/r02a_tf32_peak.err
N=8 python tools/parity_repeat.py > gpurun_out/r02a_repeat.log 2>&1
TN_TC_3XTF32=1 N=8 python tools/parity_repeat.py >> gpurun_out/r02a_repeat.log 2>&1
timeout 600 python bench.py --model m --blocks 10 --loss arc --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_m_arc.json 2> gpurun_out/r02a_m_arc.err
timeout 600 python bench.py --model l --blocks 5 --loss arc --ragged --seconds 8 --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_l_arc_ragged.json 2> gpurun_out/r02a_l_arc_ragged.err
tail -3 gpurun_out/r02a_repeat.log; cat gpurun_out/r02a_tf32_peak.json; head -c 600 gpurun_out/r02a_m_arc.json; echo; head -c 600 gpurun_out/r02a_l_arc_ragged.json; tail -3 gpurun_out/r02a_m_arc.err gpurun_out/r02a_l_arc_ragged.err
