#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/multi_pytest.log 2>&1
tail -25 gpurun_out/multi_pytest.log | cut -c1-200
