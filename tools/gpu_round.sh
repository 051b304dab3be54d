mkdir -p gpurun_out; rm -f gpurun_out/attn_test.log
timeout 240 python -m pytest tests/test_attn_gpu.py -q -s > gpurun_out/attn_pytest.log 2>&1; A=$?
echo "attn rc=$A"; grep -E "per call|passed|failed|Error|error" gpurun_out/attn_pytest.log | tail -12
if [ $A -ne 0 ]; then export PXR_FUSED_ATTN=0; echo "FALLBACK: PXR_FUSED_ATTN=0"; tail -30 gpurun_out/attn_pytest.log; fi
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_attn_gpu.py > gpurun_out/gpu_pytest.log 2>&1; G=$?
echo "gpu tests rc=$G"; tail -6 gpurun_out/gpu_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench new rc=$?"; cat gpurun_out/bench_new.json | cut -c1-200
timeout 300 python tools/profile_ops.py gpurun_out/ops_new.csv > gpurun_out/ops_new.log 2>&1; head -12 gpurun_out/ops_new.csv
