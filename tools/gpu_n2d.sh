#!/bin/bash
# 2-GPU box: BFS tests on one GPU, the multi-fragment tests, N=1 BFS bench, N=2 bench + trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_apps.py tests/test_gpu_multi.py tests/test_gpu_engine.py -q -m gpu -x -k "bfs or multi or engine" > gpurun_out/n2_pytest.log 2>&1
tail -4 gpurun_out/n2_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sweep none > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n1_bench.json').read().strip().splitlines()[-1])
print('N=1 bench: bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], d['config'].get('parity'))
PY
N=2
GL_KTIME=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --sweep ${SWEEP:-none} > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo bench rc=$?
grep gl-ktime gpurun_out/n${N}_bench.err | tail -4
bash tools/gpu_n2trace.sh 2 | grep -v "k_\|last xsync"
python - <<PY
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print('N=2 bench: bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], 'setup', d['config'].get('app_setup_ms'), d['config'].get('parity'))
PY
