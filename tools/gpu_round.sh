#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_image_prompts_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/last_check.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/last_check.log
