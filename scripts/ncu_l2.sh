#!/bin/bash
# usage (on the GPU box): scripts/ncu_l2.sh <variant-name|default> <pairs>  -> gpurun_out/ncu_l2_<name>.csv
NAME=$1; PAIRS=${2:-16}
[ "$NAME" != "default" ] && export B2OF_LIB=$PWD/hackathonopticalflow_b200/csrc/variants/libb2of_$NAME.so
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_op_read_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:fb_iter --launch-skip 12 --launch-count 12 --csv --log-file gpurun_out/ncu_l2_$NAME.csv python scripts/gpu_fb_once.py $PAIRS > gpurun_out/ncu_l2_$NAME.log 2>&1
