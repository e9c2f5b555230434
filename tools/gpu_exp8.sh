#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/exp8_pytest.log 2>&1
tail -4 gpurun_out/exp8_pytest.log
timeout 900 python bench.py > gpurun_out/exp8_bench.json 2> gpurun_out/exp8_bench.err
echo bench rc=$?
python tools/ref_gpu_compare.py --scale 22 --repeat 3 > gpurun_out/exp8_refgpu_s22.log 2>&1
tail -48 gpurun_out/exp8_refgpu_s22.log
