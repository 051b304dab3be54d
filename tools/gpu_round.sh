mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_aux_losses_gpu.py -q -s > gpurun_out/aux_pytest.log 2>&1; echo "aux rc=$?"; grep -E "^\[aux\]|passed|failed|Error|assert" gpurun_out/aux_pytest.log | head -40
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_aux_losses_gpu.py > gpurun_out/gpu_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_pytest.log
timeout 300 python tools/profile_ops.py gpurun_out/ops_new.csv > gpurun_out/ops_new.log 2>&1; grep -v "^gemm\|^conv" gpurun_out/ops_new.csv | head -40
PXR_GN_COOP=0 timeout 300 python tools/profile_ops.py gpurun_out/ops_gn3k.csv > gpurun_out/ops_gn3k.log 2>&1; grep "^gn_" gpurun_out/ops_gn3k.csv | head -30
