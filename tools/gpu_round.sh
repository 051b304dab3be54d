mkdir -p gpurun_out
timeout 900 python tools/vdiff_layers.py 64 2>&1 | grep -v Warning > gpurun_out/vdiff_layers.log; tail -150 gpurun_out/vdiff_layers.log
