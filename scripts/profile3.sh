#!/bin/bash
# one ncu --set full capture of the dominant kernels of the headline step + the launch list of the same command
TAG=${1:-r1c}
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"k_lz4_frames|k_encode_str|k_str_sizes|k_encode_fixed" -s 12 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --cpu-budget 0.1 --no-extra > gpurun_out/prof_$TAG.log 2>&1 < /dev/null
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --cpu-budget 0.1 --no-extra > gpurun_out/launches_$TAG.log 2>&1 < /dev/null
ls -la gpurun_out/ | tail -6
