"""Extracts, from the reference's own emitter fixtures, what pins the Debezium emitter restatement (oracle/oracle.cpp
orc_debezium_emit) for columns whose database-specific converter is the identity:

  pkg/debezium/pg/tests/testdata/emitter_crud_test__insert.txt           the ChangeItem the reference's test emits
  pkg/debezium/pg/tests/testdata/emitter_crud_test__debezium_insert.txt  the message a real Debezium wrote for that row
                                                                          (emitter_crud_test.go compares against it)

Run in the build container (reads /root/reference): python tests/golden/make_debezium_emit_goldens.py
"""
import json
import os
import re

REF = "/root/reference/pkg/debezium/pg/tests/testdata"
IDENTITY = re.compile(r"^pg:(integer|smallint|bigint|boolean|text|character varying\(\d+\)|double precision)$")


def main():
    item = json.load(open(os.path.join(REF, "emitter_crud_test__insert.txt")))
    raw = open(os.path.join(REF, "emitter_crud_test__debezium_insert.txt"), encoding="utf-8").read()
    msg = json.loads(raw)
    cols = []
    for name, val in zip(item["columnnames"], item["columnvalues"]):
        sch = next(c for c in item["table_schema"] if c["name"] == name)
        if not IDENTITY.match(sch["original_type"]):
            continue
        # the exact text of the value in the real message's `after` object
        m = re.search(r'"after":\{.*?"%s":("(?:[^"\\]|\\.)*"|[^,}]+)' % re.escape(name), raw)
        cols.append({"name": name, "type": sch["type"], "key": bool(sch["key"]), "required": bool(sch["required"]), "value": val, "after_text": m.group(1)})
    out = {
        "source": "pkg/debezium/pg/tests/testdata/emitter_crud_test__insert.txt + emitter_crud_test__debezium_insert.txt",
        "table": [item["schema"], item["table"]], "id": item["id"], "lsn": item["nextlsn"], "commit_time": item["commitTime"],
        "columns": cols,
        "payload_keys": sorted(msg["payload"].keys()), "source_block": msg["payload"]["source"], "op": msg["payload"]["op"],
        # assertions of the reference's own unit tests on the common path (emitter_value_converter_test.go:62-79,
        # mysql/tests/emitter_meta_test.go:15-70)
        "substrings": {"html": '"value":"<>!@#$%^&*()_"', "mysql_file": '"file":"mysql-log.000002"', "mysql_pos": '"pos":13747',
                       "mysql_gtid": '"gtid":"58c4f6fc-27b5-11ed-b434-0242ac1e0002:2"'},
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "debezium_emit_goldens.json")
    json.dump(out, open(dst, "w", encoding="utf-8"), ensure_ascii=False, indent=1)
    print(len(cols), "identity columns ->", dst)


if __name__ == "__main__":
    main()
