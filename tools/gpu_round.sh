mkdir -p gpurun_out
timeout 300 python tools/gemm_square.py 2>&1 | tail -5
timeout 600 python bench.py --steps 30 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 760 -c 400 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_c2.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/ncu_launches.log
timeout 300 python tools/profile_ops.py gpurun_out/ops_new.csv > gpurun_out/ops_new.log 2>&1; head -5 gpurun_out/ops_new.csv
