mkdir -p gpurun_out
timeout 900 python tools/bench_vdiff.py 10 2>&1 | grep -E "setup|workload|Error|error" | cut -c1-700
