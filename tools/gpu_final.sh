#!/bin/bash
# 1-GPU box: the whole -m gpu suite, the default bench line, a 2-fragment bench line on one device
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_pytest.log 2>&1
tail -3 gpurun_out/final_pytest.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo bench rc=$?
(GL_ONE_DEVICE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --scale 20 --steps 3 --warmup 3 --no-cpu-baseline --sweep none > gpurun_out/final_onedev2.json 2> gpurun_out/final_onedev2.err; echo onedev rc=$?)
python - <<PY
import json
for f in ('final_bench','final_onedev2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    c=d['config']; apps=c.pop('apps',None) or {}
    print(f,'ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['alg_bytes_per_launch'],'parity',c.get('parity',{}).get('parity_ok'),'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'))
    print('   ', c['ms_per_superstep'], c.get('superstep_mode'))
    for k,v in apps.items():
        print('   ',k,{kk:v[kk] for kk in ('ms_per_query','teps','supersteps','parity_ok') if kk in v})
PY
