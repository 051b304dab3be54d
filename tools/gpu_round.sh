mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gpu_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench new rc=$?"; cut -c1-200 gpurun_out/bench_new.json
PXR_GEMM_CTA_GROUP=2 timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_cg2.json 2> /dev/null; echo "bench cg2:"; cut -c1-200 gpurun_out/bench_cg2.json
timeout 300 python tools/profile_ops.py gpurun_out/ops_new.csv > gpurun_out/ops_new.log 2>&1; grep "^gemm M=12608\|^attn\|^conv3x3 M=65536" gpurun_out/ops_new.csv | head -20
PXR_GEMM_CTA_GROUP=2 timeout 300 python tools/profile_ops.py gpurun_out/ops_cg2.csv > gpurun_out/ops_cg2.log 2>&1; grep "^gemm M=12608\|^conv3x3 M=65536" gpurun_out/ops_cg2.csv | head -20
