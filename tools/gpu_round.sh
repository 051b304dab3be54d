#!/bin/bash
# scratch driver for one gpurun call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_color_jitter.py tests/test_image_prompts_gpu.py -m gpu -q -s > gpurun_out/jitter_imgprompt_pytest.log 2>&1; echo "jitter+image prompts rc=$?"
grep -E "parity|passed|failed|Error|assert|error" gpurun_out/jitter_imgprompt_pytest.log | tail -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cutout_ -c 8 --csv --log-file gpurun_out/ncu_cutout_launches.csv python tools/profile_c2.py 3 > gpurun_out/ncu_cutout.log 2>&1; echo "ncu rc=$?"
grep -E "cutout_(fwd|bwd)" gpurun_out/ncu_cutout_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
timeout 900 python bench.py --steps 30 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_full.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_full.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench_full.json | cut -c1-400; tail -3 gpurun_out/bench_full.err
