#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
GL_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --sweep none > gpurun_out/n${N}_trace.json 2> gpurun_out/n${N}_trace.err
grep 'gl-trace\]' gpurun_out/n${N}_trace.err | tail -24
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_trace.json').read().strip().splitlines()[-1])
print('N=${N} bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12, d['config'].get('ms_per_superstep'))
PY
