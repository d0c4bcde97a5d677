#!/bin/bash
# round-end style run on one GPU: full GPU suite, smoke, bench (+ reference arm), ncu captures. Every step is hard-killed on timeout.
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests -q -m gpu 2>&1 < /dev/null | tail -8 | cut -c1-400
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 < /dev/null | tail -2
timeout -s KILL 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err < /dev/null; tail -c 300 gpurun_out/bench_$TAG.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err < /dev/null
bash scripts/profile3.sh $TAG 2>&1 | tail -3
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g frac %.4f launches %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print(d["roofline"]["all_kernels_ms"])
print({k: (round(v["rows_per_s"]), round(v["ms"],1)) for k,v in d.get("other_paths",{}).items() if isinstance(v,dict)})
print(open("gpurun_out/bench_ref_$TAG.json").read()[:300])
PY
