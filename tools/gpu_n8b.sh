#!/bin/bash
# 8-GPU box: the N=8 bench line (with the C4 / C5 sweep and the in-run parity checks) and a phase trace
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
GL_KTIME=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo bench rc=$?
grep gl-ktime gpurun_out/n${N}_bench.err | grep kernel | tail -3
GL_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --sweep none --no-parity > gpurun_out/n${N}_trace.json 2> gpurun_out/n${N}_trace.err
grep 'gl-trace\] f0 level' gpurun_out/n${N}_trace.err | tail -8
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
apps=d['config'].pop('apps',{}) or {}
print('N=${N} bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], d['config'].get('parity'))
for k,v in apps.items():
    print(k, {kk:v[kk] for kk in ('ms_per_query','teps','supersteps','parity_ok') if kk in v})
PY
