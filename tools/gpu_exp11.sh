#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_apps.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -m gpu -k "bfs or multi" > gpurun_out/exp11_pytest.log 2>&1
tail -3 gpurun_out/exp11_pytest.log
for i in 1 2 3; do python bench.py --no-cpu-baseline --sweep none --steps 10 --warmup 3 > gpurun_out/exp11_bfs_$i.json 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp11_bfs_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms/query %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], d['config']['superstep_mode'])
PY
