#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_compat.py -q -m gpu -k "lcc or engine or batch or wcc_opt or multi" > gpurun_out/exp6_pytest.log 2>&1
tail -3 gpurun_out/exp6_pytest.log
B="python bench.py --no-cpu-baseline --sweep none --steps 5 --warmup 2"
$B > gpurun_out/exp6_bfs.json 2>&1
GL_HOST_THREADS=48 $B > gpurun_out/exp6_bfs_t48.json 2>&1
GL_HOST_THREADS=16 $B > gpurun_out/exp6_bfs_t16.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp6_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('exp6_')[1], 'ms/query %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', open(f).read()[-300:])
PY
