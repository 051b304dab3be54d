mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/gpu_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench new rc=$?"; cut -c1-200 gpurun_out/bench_new.json
PXR_CONV_SPLITK=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_nosplitk.json 2> /dev/null; echo "bench no-splitk:"; cut -c1-200 gpurun_out/bench_nosplitk.json
timeout 300 python tools/profile_ops.py gpurun_out/ops_new.csv > gpurun_out/ops_new.log 2>&1; grep -v "^gemm M=12608" gpurun_out/ops_new.csv | head -60
