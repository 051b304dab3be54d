#!/bin/bash
# scratch driver for one gpurun call (2 GPUs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -s -k "image_prompts or aux" > gpurun_out/multigpu_pytest.log 2>&1; echo "multigpu rc=$?"
grep -E "parity|passed|failed|Error|assert|error" gpurun_out/multigpu_pytest.log | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/bench_2gpu.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_2gpu.json; tail -2 gpurun_out/bench_2gpu.err
