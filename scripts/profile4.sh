#!/bin/bash
# one ncu --set full capture of the dominant kernels of the headline step + the launch list of the same command
TAG=${1:-r2a}; KERN=${2:-k_lz4_frames|k_frame_seal}
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/lz4_phases.py > gpurun_out/phases_$TAG.txt 2>&1 < /dev/null
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s 6 -c 2 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --cpu-budget 0.1 --no-extra > gpurun_out/prof_$TAG.log 2>&1 < /dev/null
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --cpu-budget 0.1 --no-extra > gpurun_out/launches_$TAG.log 2>&1 < /dev/null
tail -3 gpurun_out/phases_$TAG.txt
ls -la gpurun_out/ | tail -6
