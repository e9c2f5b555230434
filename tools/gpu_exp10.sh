#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/exp10_pytest.log 2>&1
tail -3 gpurun_out/exp10_pytest.log
B="python bench.py --no-cpu-baseline --steps 5 --warmup 2"
$B > gpurun_out/exp10_bench.json 2> gpurun_out/exp10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/exp10_bench.json').read().strip().splitlines()[-1])
apps=d['config'].pop('apps')
print('bfs',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
for k,v in apps.items():
    print(k, {kk:(round(v[kk],4) if isinstance(v[kk],float) else v[kk]) for kk in ('ms_per_query','ms_per_round','frac_whole_query','parity_ok') if kk in v}, v.get('pull_variant',{}).get('ms_per_round'))
PY
