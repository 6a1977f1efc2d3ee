#!/bin/bash
# one GPU-box pass for the round's evidence: parity tests, default bench, ncu launch list of the same bench
# command, one full capture of the dominant kernel (finest-level fb_iter_ws, flow-in mode)
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v12.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fb_iter_ws --launch-skip 22 --launch-count 1 -f -o gpurun_out/prof_v12_ws python scripts/gpu_fb_once.py 16 > gpurun_out/prof_v12_ws.log 2>&1
tail -2 gpurun_out/prof_v12_ws.log
