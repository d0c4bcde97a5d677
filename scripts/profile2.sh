#!/bin/bash
TAG=${1:-r1b}; KERN=${2:-k_lz4_frames|k_frame_seal|k_encode_str}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s 9 -c 3 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --cpu-budget 0.1 > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/ | tail -5
