"""Extract the generic JSON parser goldens from the reference tree into tests/golden/json_parser_goldens.json.
Run in the build container only (needs /root/reference):  python tests/golden/make_json_parser_goldens.py
Sources: pkg/parsers/generic/test_data/*.jsonl (inputs, parser_test.go:19-29) and
         pkg/parsers/generic/gotest/canondata/result.json (canonised ChangeItems of TestParserNumberTypes, TestBase64Unpack)."""
import json, os
R = "/root/reference/pkg/parsers/generic"
canon = json.load(open(f"{R}/gotest/canondata/result.json"))   # note: the canon file was normalised by a Python tool (1e-07, 100000.0): compare VALUES, not spelling
out = {"inputs": {n: open(f"{R}/test_data/{n}").read() for n in ("parser_numbers_test.jsonl", "parse_base64_packed.jsonl", "parser_unescape_test.jsonl")}, "canon": {}}
for k, v in canon.items():
    name = k.split(".")[-1]
    def slim(items):
        return [{"columnnames": it["columnnames"], "columnvalues": it["columnvalues"], "table": it["table"], "nextlsn": it["nextlsn"],
                 "types": [c["type"] for c in it["table_schema"]]} for it in items]
    out["canon"][name] = {m: slim(x) for m, x in v.items()} if isinstance(v, dict) else slim(v)
# pkg/parsers/tests/generic_parser_test.go:274-296 TestParser_DoJson: the sample, its field list, and what the test asserts (36 rows,
# column `version` = 89488198116272410 + row: uint64 values beyond 2^53 must arrive exactly)
S = "/root/reference/pkg/parsers/tests/samples"
cfg = json.load(open(f"{S}/json_sample.json"))["ParserConfig"]["json.lb"]
out["do_json"] = {"input": open(f"{S}/json_sample").read(), "add_rest": cfg["AddRest"], "null_keys_allowed": cfg["NullKeysAllowed"],
                  "fields": [{"name": f["name"], "type": f["type"], "key": f["key"], "required": f["required"]} for f in cfg["Fields"]],
                  "rows": 36, "version_base": 89488198116272410}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "json_parser_goldens.json"), "w"), indent=1, sort_keys=True)
