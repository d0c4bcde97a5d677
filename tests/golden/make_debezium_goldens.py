"""Debezium parser goldens from the reference tree -> tests/golden/debezium_goldens.json (run in the build container only).
Sources: pkg/parsers/registry/debezium/engine/parser_test.jsonl (input message, parser_test.go:20-27) and
         pkg/parsers/registry/debezium/engine/gotest/canondata/result.json (TestParser: the ChangeItem it must become)."""
import json, os
R = "/root/reference/pkg/parsers/registry/debezium/engine"
canon = json.load(open(f"{R}/gotest/canondata/result.json"))["gotest.gotest.TestParser"]
lines = [l for l in open(f"{R}/parser_test.jsonl").read().split("\n") if l]
out = {"messages": lines, "items": [{k: it[k] for k in ("columnnames", "columnvalues", "commitTime", "id", "kind", "nextlsn", "schema", "table")} | {"types": [(c["name"], c["type"], c["key"]) for c in it["table_schema"]]} for it in canon]}
# pkg/debezium/receiver_test.go:15-26 TestDelete: a delete envelope and the ChangeItem the receiver must build from it
import re
src = open("/root/reference/pkg/debezium/receiver_test.go").read()
msg = re.search(r"debeziumMsg := `(.*?)`", src, re.S).group(1)
exp = json.loads(re.search(r"require.Equal\(t, `(\{\"id\":557.*?)`, ", src, re.S).group(1))
out["delete_case"] = {"message": msg, "schema_text": msg[msg.rindex(',"schema":{"fields"') + len(',"schema":'):-1],
                      "expected": {k: exp[k] for k in ("id", "nextlsn", "kind", "schema", "table", "columnnames", "oldkeys")} |
                                  {"types": [(c["name"], c["type"], c["key"]) for c in exp["table_schema"]]}}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "debezium_goldens.json"), "w"), indent=0)
