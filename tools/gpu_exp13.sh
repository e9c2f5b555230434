#!/bin/bash
# 1-GPU box: BFS / engine / multi-fragment (one device) tests, then the N=1 BFS bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_apps.py tests/test_gpu_multi.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -q -m gpu -x -k "bfs or multi or engine" > gpurun_out/exp13_pytest.log 2>&1
tail -4 gpurun_out/exp13_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sweep none > gpurun_out/exp13_bench.json 2> gpurun_out/exp13_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/exp13_bench.json').read().strip().splitlines()[-1])
print('N=1 bench: bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], d['config'].get('parity',{}).get('parity_ok'))
PY
done
