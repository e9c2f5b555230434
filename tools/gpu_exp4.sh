#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/exp4_pytest.log 2>&1
echo rc=$? >> gpurun_out/exp4_pytest.log
tail -5 gpurun_out/exp4_pytest.log
(GL_ONE_DEVICE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --scale 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/exp4_bench_onedev2.json 2> gpurun_out/exp4_bench_onedev2.err; echo onedev rc=$?)
B="python bench.py --no-cpu-baseline --sweep none --steps 5 --warmup 2"
for v in 1 0; do
  GL_HUB_TMA=$v $B > gpurun_out/exp4_bfs_tma$v.json 2>&1
  GL_HUB_TMA=$v $B --push-only --no-fuse > gpurun_out/exp4_bfs_push_nofuse_tma$v.json 2>&1
  GL_HUB_TMA=$v $B --app pagerank > gpurun_out/exp4_pr_push_tma$v.json 2>&1
  GL_HUB_TMA=$v $B --app sssp > gpurun_out/exp4_sssp_tma$v.json 2>&1
  GL_HUB_TMA=$v $B --app wcc > gpurun_out/exp4_wcc_tma$v.json 2>&1
done
$B --app pagerank --pr-pull > gpurun_out/exp4_pr_pull64.json 2>&1
$B --app pagerank --pr-f32 > gpurun_out/exp4_pr_pull32_hub.json 2>&1
$B --app pagerank --pr-f32 --pr-nohub > gpurun_out/exp4_pr_pull32_nohub.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('exp4_')[1], 'ms/query %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['kernel'][:40])
    except Exception as e:
        print(f, 'ERR', open(f).read()[-300:])
PY
