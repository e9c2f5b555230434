#!/bin/bash
# 4-GPU box: the N=4 bench line (BFS + in-run parity check; C4 / C5 sweep skipped to keep the call short)
N=4
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --sweep none > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/n4_bench.json').read().strip().splitlines()[-1])
print('N=4 bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], d['config'].get('parity'), 'frac', d['roofline']['frac'])
PY
