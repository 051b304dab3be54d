mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vdiff_gpu.py -q -s -x -k 64 > gpurun_out/vdiff_pytest.log 2>&1; echo "vdiff rc=$?"; grep -E "parity|passed|failed|Error|error|assert" gpurun_out/vdiff_pytest.log | head -30; tail -5 gpurun_out/vdiff_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 900 -c 470 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_c2.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/ncu_launches.log
