"""Profiling aid (GPU): one fused CSV -> native+LZ4 call over the hits-shaped CSV of bench.py's config #5 leg."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transferia_b200 import abi, engine, workload
import bench
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cb, cschema = bench.make_batch(rows, workload.SEED)
cschema = [dict(c, path=str(i)) for i, c in enumerate(cschema)]
ctext = workload.render_hits_csv(cb, cschema)
eng = engine.Engine(0)
pid = eng.plan("public", "hits", cschema, [], {"type": "clickhouse"})
for _ in range(3):
    r, _c = eng.parse_csv(pid, ctext, wire_fmt=abi.TF_WIRE_CH_NATIVE_LZ4)
print(r.rows_out, len(r.wire))
