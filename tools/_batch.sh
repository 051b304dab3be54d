set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bookkeeping_gpu.py tests/test_zz_host_paths_gpu.py tests/test_gemm_gpu.py -m gpu -x -q -k "not throughput" > gpurun_out/r02_b2_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_b2_tests.log
timeout 600 python tests/diag_color_jitter.py > gpurun_out/r02_diag_jitter.log 2>&1; echo "diag rc=$?"
timeout 1200 python -m pytest tests/test_big_shapes_gpu.py -m gpu -s -q > gpurun_out/r02_b2_bigshapes.log 2>&1; echo "big rc=$?"
tail -5 gpurun_out/r02_b2_bigshapes.log
PXR_GN_COOP=0 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02_sanitizer_memcheck.log
PXR_GN_COOP=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r02_sanitizer_racecheck.log
