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


# original types the device emitter handles (pkg/debezium/pg/emitter.go:265-629), with the cell the strict columnar layout holds
PG_FORMS = {
    "pg:boolean": "bool", "pg:smallint": "int", "pg:integer": "int", "pg:bigint": "int", "pg:real": "f64", "pg:double precision": "f64",
    "pg:text": "str", "pg:uuid": "str", "pg:cidr": "str", "pg:macaddr": "str", "pg:citext": "str", "pg:int4range": "str", "pg:int8range": "str", "pg:inet": "str",
    "pg:bytea": "b64", "pg:json": "json", "pg:jsonb": "json", "pg:date": "time", "pg:timestamp with time zone": "time",
}


def pg_cell(otype, val):
    """(kind, payload) of the typed cell for a canon ChangeItem value, or None when the device does not take the type."""
    import datetime
    kind = PG_FORMS.get(otype)
    if kind is None and re.match(r"^pg:character( varying)?(\(\d+\))?$", otype):
        kind = "str"
    wall = False
    if kind is None and re.match(r"^pg:timestamp(\(\d\))? without time zone$", otype):
        kind, wall = "time", True
    if kind is None:
        return None
    if kind == "time":
        t = datetime.datetime.fromisoformat(val.replace("Z", "+00:00"))
        if wall:        # pgtype.Timestamp.Set keeps the wall-clock fields: a UTC-only layout holds that wall clock as the UTC instant
            t = t.replace(tzinfo=datetime.timezone.utc)
        d = t - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
        return "time", [d.days * 86400 + d.seconds, d.microseconds * 1000]
    if kind == "json":
        return "json", json.dumps(val, separators=(",", ":"), sort_keys=True)
    return kind, val


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
    pg = []
    for name, val in zip(item["columnnames"], item["columnvalues"]):
        sch = next(c for c in item["table_schema"] if c["name"] == name)
        cell = pg_cell(sch["original_type"], val)
        if cell is None:
            continue
        pg.append({"name": name, "type": sch["type"], "original_type": sch["original_type"], "key": bool(sch["key"]), "required": bool(sch["required"]),
                   "kind": cell[0], "cell": cell[1], "after": msg["payload"]["after"][name]})
    out = {
        "pg_columns": pg,
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
    print(len(cols), "identity columns,", len(pg), "pg-typed columns ->", dst)


if __name__ == "__main__":
    main()
