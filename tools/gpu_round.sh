mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vdiff_gpu.py -q -s -k "64 or matched" > gpurun_out/vdiff_pytest.log 2>&1; echo "vdiff rc=$?"; grep -E "parity|passed|failed|Error|error" gpurun_out/vdiff_pytest.log | head -30
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_vdiff_gpu.py > gpurun_out/gpu_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_new.json
PXR_GN_CLUSTER=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_nocluster.json 2> /dev/null; echo "bench no-cluster:"; cut -c1-200 gpurun_out/bench_nocluster.json
