#!/bin/bash
mkdir -p gpurun_out
python tools/ref_gpu_compare.py --scale 22 --repeat 3 > gpurun_out/exp7_refgpu_s22.log 2>&1
tail -45 gpurun_out/exp7_refgpu_s22.log
