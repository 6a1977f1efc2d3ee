#!/bin/bash
# usage (GPU box): scripts/ncu_full.sh <variant|default> <kernel-regex> <skip> <out-name> [pairs]
NAME=$1; K=$2; SKIP=$3; OUT=$4; PAIRS=${5:-16}
[ "$NAME" != "default" ] && export B2OF_LIB=$PWD/hackathonopticalflow_b200/csrc/variants/libb2of_$NAME.so
ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip $SKIP --launch-count 1 -f -o gpurun_out/$OUT python scripts/gpu_fb_once.py $PAIRS > gpurun_out/$OUT.log 2>&1
