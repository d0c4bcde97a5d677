"""Profiling aid (GPU): the JSON-lines path of bench.py's config #2 leg (parse -> mask_field -> ClickHouse JSONEachRow), a few calls."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transferia_b200 import abi, engine, workload
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
text, fields = workload.make_json_lines(lines)
opts = {"add_rest": True, "add_dedupe_keys": True, "partition": '{"partition":0,"topic":"events"}'}
schema = engine.json_result_schema(fields, opts)
trs = [{"mask_field": {"columns": ["user"], "maskFunctionHash": {"userDefinedSalt": "pepper"}}}]
eng = engine.Engine(0)
pid = eng.plan("", "events", schema, trs, {"type": "clickhouse"})
for _ in range(3):
    r = eng.parse_json(pid, text, opts, None, wire_fmt=abi.TF_WIRE_CH_JSONEACHROW)
print(r.rows_out, len(r.wire))
