#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/exp12_pytest.log 2>&1
tail -4 gpurun_out/exp12_pytest.log
B="python bench.py --no-cpu-baseline --sweep none --steps 5 --warmup 2"
for a in bfs sssp pagerank wcc; do $B --app $a > gpurun_out/exp12_$a.json 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp12_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('exp12_')[1], 'ms/query %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], [round(x,3) for x in d['config']['ms_per_superstep'][:8]])
    except Exception as e:
        print(f, 'ERR', open(f).read()[-300:])
PY
