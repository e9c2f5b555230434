#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/exp3_pytest.log 2>&1
echo rc=$? >> gpurun_out/exp3_pytest.log
tail -5 gpurun_out/exp3_pytest.log
timeout 900 python bench.py > gpurun_out/exp3_bench.json 2> gpurun_out/exp3_bench.err
echo bench rc=$?
echo skip ref
echo ref rc=$?
free -g | head -2; nproc
