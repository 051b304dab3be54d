mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vdiff_gpu.py -q -s -k "64" 2>&1 | grep -E "z.grad|passed|failed|Error" | head
timeout 900 python tools/bench_vdiff.py 10 2>&1 | grep -E "workload|events" | cut -c1-330
grep -E "^gemm M=16|^vd_gn_fwd px=16 |^gemm M=64" gpurun_out/vdiff_ops.csv | head
