for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default        ', d['ms_per_step'], d['value'])"
TN_FUSE_DWFWD=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TN_FUSE_DWFWD=1', d['ms_per_step'], d['value'])"
done
TN_FUSE_DWFWD=1 python tools/graph_profile.py 2>/dev/null | grep -E "gemm_tc2|dw_fwd|sum of"
