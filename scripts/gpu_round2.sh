#!/bin/bash
# round-2 evidence pass on one GPU box: default bench, reference arm, ncu launch list of the same bench command (chunk
# on one stream, B2OF_STREAMS=1, so that the list is the per-launch view kernel_ms_per_step reports), full captures
# (with source) of the dominant kernel, of the per-frame / upsample kernels and of lk_track
set -x
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1; tail -c 300 gpurun_out/r02_bench_reference.json
B2OF_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
B2OF_STREAMS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fb_iter_ws --launch-skip 22 --launch-count 1 -f -o gpurun_out/prof_r2_ws python scripts/gpu_fb_once.py 16 > gpurun_out/prof_r2_ws.log 2>&1
B2OF_STREAMS=1 timeout 600 ncu --set full --clock-control none -k "regex:fb_levels_coarse|fb_upsample2x|fb_level0_stream|fb_polyexp" --launch-skip 9 --launch-count 9 -f -o gpurun_out/prof_r2_frame python scripts/gpu_fb_once.py 64 > gpurun_out/prof_r2_frame.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lk_track --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r2_lk python scripts/gpu_lk_profile.py > gpurun_out/prof_r2_lk.log 2>&1
for f in ws frame lk; do tail -n 2 gpurun_out/prof_r2_$f.log; done
