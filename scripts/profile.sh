#!/bin/bash
# Run under gpurun (1 GPU). Produces gpurun_out/{launches,prof}_$TAG.* ; summaries are copied to profiles/ by hand.
TAG=${1:-r1}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 44 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --cpu-budget 0.2 > gpurun_out/launches_$TAG.log 2>&1
# the heavy kernels, full sections + source
ncu --set full --clock-control none --import-source on -k regex:'k_lz4_frames|k_frame_seal|k_encode_str|k_encode_fixed' -s 12 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.2 > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/
