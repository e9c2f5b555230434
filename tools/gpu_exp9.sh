#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -q -m gpu > gpurun_out/exp9_pytest.log 2>&1
tail -30 gpurun_out/exp9_pytest.log | grep -E 'mgpu|passed|failed|Error|error' | tail -40
B="python bench.py --no-cpu-baseline --sweep none --steps 5 --warmup 2 --scale 22"
for a in bfs sssp pagerank wcc; do $B --app $a > gpurun_out/exp9_s22_$a.json 2>&1; done
$B --app wcc --wcc-opt > gpurun_out/exp9_s22_wccopt.json 2>&1
$B --app pagerank --pr-f32 > gpurun_out/exp9_s22_prhub.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp9_s22_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('exp9_')[1], 'ms/query %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', open(f).read()[-300:])
PY
