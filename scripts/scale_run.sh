#!/bin/bash
# bench.py at the GPU counts given as arguments, back to back on one box (run under gpurun --gpus N): one JSON line per N into gpurun_out/scale_<tag>.jsonl
TAG=${TAG:-r2}
mkdir -p gpurun_out
: > gpurun_out/scale_$TAG.jsonl
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --cpu-budget 1 2>> gpurun_out/scale_$TAG.err | tail -1 >> gpurun_out/scale_$TAG.jsonl
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra 2>> gpurun_out/scale_$TAG.err | grep '^{' | tail -1 >> gpurun_out/scale_$TAG.jsonl
  fi
done
python - <<'PY'
import json
for l in open("gpurun_out/scale_%s.jsonl" % __import__("os").environ.get("TAG", "r2")):
    d = json.loads(l); print(d["n_gpus"], "GPUs: value %.4g rows/s (%.3f ms/step), e2e %.4g rows/s" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
