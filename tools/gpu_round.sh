#!/bin/bash
# scratch driver for one gpurun call: the bench exactly as the driver launches it
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_default.json)"
cut -c1-260 gpurun_out/bench_default.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_default.json | cut -c1-200; tail -2 gpurun_out/bench_default.err
