#!/bin/bash
# closing evidence pass of round 2 (kernels changed since gpu_round2.sh ran: lk_track, gftt_select): default bench line,
# reference arm, ncu launch list of the same bench command (one stream), full captures of lk_track and the GFTT kernels
set -x
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1; tail -c 300 gpurun_out/r02_bench_reference.json
B2OF_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lk_track --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r2_lk python scripts/gpu_lk_profile.py > gpurun_out/prof_r2_lk.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:gftt_" --launch-skip 6 --launch-count 3 -f -o gpurun_out/prof_r2_gftt python scripts/gpu_gftt_profile.py > gpurun_out/prof_r2_gftt.log 2>&1
for f in lk gftt; do tail -n 2 gpurun_out/prof_r2_$f.log; done
python scripts/gpu_call_latency.py > gpurun_out/r02_call_latency.txt 2>&1; tail -5 gpurun_out/r02_call_latency.txt
