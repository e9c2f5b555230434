#!/bin/bash
# N=2 box: multi-fragment parity tests (real 2 GPUs + one-device mode), then the N=2 bench line and a phase trace
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/n2_pytest.log 2>&1
tail -5 gpurun_out/n2_pytest.log
N=2
GL_KTIME=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --sweep ${SWEEP:-none} > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo bench rc=$?
grep gl-ktime gpurun_out/n${N}_bench.err | tail -6
bash tools/gpu_n2trace.sh 2
python - <<PY
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print('N=2 bench: bfs ms',d['ms_per_step'],'TTEPS',d['value']/1e12,'e2e ms',d['e2e']['ms_per_step'], d['config']['ms_per_superstep'], 'setup', d['config'].get('app_setup_ms'))
PY
