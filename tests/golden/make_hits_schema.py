"""Extracts the ClickBench `hits` table schema (first 99 columns) and the per-column
fixture cell from the reference's own fixture
    /root/reference/pkg/providers/postgres/testdata/hits_data.json  (keys parse_schema, data)
into tests/golden/hits_schema.json.  Run in the build container only (the reference tree does
not exist on the GPU box); the output is committed.
"""
import base64, json, os
src = "/root/reference/pkg/providers/postgres/testdata/hits_data.json"
d = json.load(open(src))
cols = []
row = d["data"][0]
for c, cell in list(zip(d["parse_schema"], row))[:99]:
    raw = base64.b64decode(cell) if cell else b""
    cols.append({
        "table_schema": c["table_schema"], "table_name": c["table_name"], "path": c["path"], "name": c["name"],
        "type": c["type"], "key": c["key"], "fake_key": c["fake_key"], "required": c["required"],
        "expression": c["expression"], "original_type": c["original_type"],
        "fixture_cell_b64": base64.b64encode(raw).decode(),
    })
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hits_schema.json")
json.dump({"source": "pkg/providers/postgres/testdata/hits_data.json (reference commit 6affc0ca)", "columns": cols}, open(out, "w"), indent=1)
print(len(cols), "columns ->", out)
