"""Debezium parser: the oracle against the reference's canon ChangeItem on CPU; the device parser against the
oracle on GPU."""
import base64
import json
import os

import numpy as np
import pytest

from transferia_b200 import abi

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "debezium_goldens.json")))


def cell(batch, c, r):
    col = batch.columns[c]
    if col.validity is not None and not (col.validity[r >> 3] >> (r & 7)) & 1:
        return None
    if col.type in abi.VAR_TYPES:
        return bytes(col.heap[col.offsets[r]:col.offsets[r + 1]])
    return col.values[r].item()


def test_parser_canon(po):
    """TestParser (parser_test.go:42-57): the pg basic_types message -> the canonised ChangeItem, column by column."""
    msg = G["messages"][0].encode(); it = G["items"][0]
    schema_text = json.dumps(json.loads(msg)["schema"], separators=(",", ":"))
    # the exact bytes of the embedded schema are what the plan is keyed on
    raw = msg.decode(); a = raw.index('"schema":') + len('"schema":'); b = raw.index(',"payload":'); schema_text = raw[a:b]
    batch, kinds, tx, lsn, ct, rm, errs, schema = po.debezium_parse(msg, [len(msg)], schema_text)
    assert errs == [] and batch.nrows == 1
    assert [(c["name"], c["type"], c["key"]) for c in schema] == [tuple(t) for t in it["types"]]
    assert kinds[0] == abi.TF_KIND_INSERT and tx[0] == it["id"] and lsn[0] == it["nextlsn"] and ct[0] == it["commitTime"]
    for c, (name, want) in enumerate(zip(it["columnnames"], it["columnvalues"])):
        got = cell(batch, c, 0); ty = schema[c]["type"]
        if want is None: assert got is None, name
        elif ty == "string": assert got == base64.b64decode(want), name          # []byte cells are printed as base64 in the canon file
        elif ty == "utf8": assert got.decode() == want, name
        elif ty == "boolean": assert bool(got) is want, name
        else: assert got == want, (name, got, want)


def test_delete_envelope(po):
    """pkg/debezium/receiver_test.go:15-26 TestDelete: op "d" -> kind delete, ID / LSN from `source`, the table schema from the
    envelope (a_id int32 key, a_name utf8), and the `before` values — what the reference stores as OldKeys {a_id: 1}; the columnar
    row carries them as the row's values with kind = delete, the shim moves the key columns into OldKeys (receiver.go:196-216)."""
    d = G["delete_case"]; exp = d["expected"]
    data = d["message"].encode()
    b, kinds, tx, lsn, ct, rm, errs, schema = po.debezium_parse(data, [len(data)], d["schema_text"])
    assert errs == [] and b.nrows == 1 and list(kinds) == [abi.TF_KIND_DELETE] and exp["kind"] == "delete"
    assert int(tx[0]) == exp["id"] and int(lsn[0]) == exp["nextlsn"] and int(ct[0]) == 1672943646565 * 10**6
    assert [[c["name"], c["type"], bool(c.get("key"))] for c in schema] == exp["types"]
    row = {c["name"]: cell(b, k, 0) for k, c in enumerate(schema)}
    assert {k: row[k] for k in exp["oldkeys"]["keynames"]} == dict(zip(exp["oldkeys"]["keynames"], exp["oldkeys"]["keyvalues"])) and row["a_name"] is None


def test_product_schema_derivation_matches_oracle(po):
    """The product's host side of tfgpu_parse_debezium (libtfgpu.so, no GPU): table schema and receivers derived from the envelope schema
    (receiver.go:46-62, receiver_engine.go:104-141) equal the oracle's for the reference's 59-column canon schema and for the Decimal /
    VariableScaleDecimal / Point / Bits shapes; schemas the device does not take are refused on the host."""
    from transferia_b200 import engine
    canon = json.dumps(json.loads(G["messages"][0])["schema"], separators=(",", ":"))
    fld = lambda typ, fname, opt=True, **kw: dict({"type": typ, "optional": opt, "field": fname}, **kw)
    env = lambda fields: json.dumps({"type": "struct", "fields": [{"type": "struct", "fields": fields, "optional": True, "field": "before"}, {"type": "struct", "fields": fields, "optional": True, "field": "after"}]})
    shapes = env([fld("int32", "id", False), fld("bytes", "dec", name="org.apache.kafka.connect.data.Decimal", parameters={"scale": "3"}), fld("bytes", "raw"),
                  fld("struct", "vsd", name="io.debezium.data.VariableScaleDecimal", fields=[]), fld("struct", "pt", name="io.debezium.data.geometry.Point", fields=[]),
                  fld("bytes", "bits", name="io.debezium.data.Bits", parameters={"length": "8"}), fld("float", "f"), fld("double", "d"), fld("int8", "i8"), fld("string", "s")])
    for text in (canon, shapes):
        got = engine.debezium_schema_validate(text)
        fields, schema = po.debezium_fields(text)
        assert [(x["name"], x["recv"], x["scale"], x["key"]) for x in got] == [tuple(f) for f in fields]
        assert [(x["name"], x["type"], x["key"]) for x in got] == [(c["name"], c["type"], bool(c.get("key"))) for c in schema]
        assert [(c["name"], c["type"], bool(c.get("key"))) for c in engine.debezium_table_schema(text)] == [(x["name"], x["type"], x["key"]) for x in got]
    for bad in (env([fld("int32", "id", __dt_original_type_info={"original_type": "pg:integer"})]), env([fld("array", "a")]), env([fld("struct", "x", name="some.other.Struct", fields=[])]),
                json.dumps({"type": "struct", "fields": [{"type": "struct", "fields": [fld("int32", "id")], "field": "after"}]}),
                json.dumps({"type": "struct", "fields": [{"type": "struct", "fields": [fld("int32", "id")], "field": "after"}, {"type": "struct", "fields": [fld("int64", "id")], "field": "before"}]}), "{not json"):
        with pytest.raises(engine.EngineError):
            engine.debezium_schema_validate(bad)


def test_base64_to_numeric(po):
    """typeutil.Base64ToNumeric (helpers.go:972-998) incl. its quirks: scale == len gives ".12", negative two's complement."""
    f = po.base64_to_numeric
    assert f("MDk=", 0) == "12345" and f("ME8=", 2) == "123.67" and f("AeJA", 0) == "123456" and f("EAAAAAAAAAAAAAAAAA==", 0) == "1267650600228229401496703205376"
    assert f(base64.b64encode(b"\xff").decode(), 0) == "-1" and f(base64.b64encode(b"\x80\x00").decode(), 1) == "-3276.8"
    assert f(base64.b64encode(b"\x0c").decode(), 2) == ".12" and f(base64.b64encode(b"\x01").decode(), 2) == "0.01" and f(base64.b64encode(b"\x00").decode(), 3) == "0"
    # pkg/debezium/typeutil/helpers_test.go: TestBase64ToNumeric (:417-421) and, read backwards, the (numeric, base64, scale) triples of
    # TestDecimalToDebezium / TestDecimalToDebeziumPrimitivesImpl (:223-254) whose text form is unambiguous
    assert f("AQ==", 2) == "0.01"
    for text, b64, scale in (("100.00", "JxA=", 2), ("-100.00", "2PA=", 2), ("2345678901", "AIvQODU=", 0), ("-2345678901", "/3Qvx8s=", 0),
                             ("1267650600228229401496703205376", "EAAAAAAAAAAAAAAAAA==", 0), ("126765060022822940149670320537.6", "EAAAAAAAAAAAAAAAAA==", 1)):
        assert f(b64, scale) == text, (b64, scale)


def test_oracle_message_shapes(po):
    """Envelope handling: unparsed / other-schema / host classes (parser.go:34-66, include_schema.go, debezium_schema.go)."""
    schema_text = '{"type":"struct","fields":[{"type":"struct","fields":[{"type":"int32","optional":false,"field":"id"},{"type":"string","optional":true,"field":"s"}],"optional":true,"field":"before"},{"type":"struct","fields":[{"type":"int32","optional":false,"field":"id"},{"type":"string","optional":true,"field":"s"}],"optional":true,"field":"after"}]}'
    src = '"source":{"lsn":7,"ts_ms":5,"txId":3,"schema":"public","table":"t","snapshot":"false"}'
    def m(payload, schema=schema_text): return ('{"schema":%s,"payload":%s}' % (schema, payload)).encode()
    msgs = [m('{"before":null,"after":{"id":1,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":2,"s":null},%s,"op":"r"}' % src), m('{"after":{"id":3,"s":"x"},%s,"op":"u"}' % src),
            m('{"before":{"id":4,"s":"d"},"after":null,%s,"op":"d"}' % src), b"", b"{}", b'{"schema":1}', m("null"), m('{"op":"c"}'), m('{"after":{"id":1},%s,"op":"c"}' % src),
            m('{"after":{"id":"1","s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1.5,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":5},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":true},%s,"op":"c"}' % src),
            m('{"after":{"id":1,"s":"a"},%s,"op":"x"}' % src), m('{"after":{"id":1,"s":"a"},%s,"op":7}' % src), m('{"after":{"id":1,"s":"a"},"source":{"lsn":-1},"op":"c"}'), m('{"after":{"id":1,"s":"a"},"source":{"txId":4294967296},"op":"c"}'),
            m('{"after":{"id":1,"s":"a"},%s,"op":"c"} trailing' % src)[:-10] + b'} x', m('{"after":{"id":1,"s":"a"},%s,"op":"c"}' % src, '{"type":"struct","fields":[]}'), m('{"after":{"id":1,"s":"__debezium_unavailable_value"},%s,"op":"c"}' % src),
            m('{"AFTER":{"id":1,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":"a\\u00e9\\ud83d\\ude00 \xff"},%s,"op":"c","extra":[1,{"a":null}]}' % src), m('{"after":{"id":9,"s":"dup"},"after":null,%s,"op":"c"}' % src),
            m('{"after":{"id":2147483648,"s":"wrap"},%s,"op":"c"}' % src), b'{"schema":%s,"payload":{"after":{"id":1,"s":"ctl\x01"},"op":"c"}}' % schema_text.encode()]
    data = b"".join(msgs); ends = np.cumsum([len(x) for x in msgs]).tolist()
    batch, kinds, tx, lsn, ct, rm, errs, schema = po.debezium_parse(data, ends, schema_text)
    assert list(rm) == [0, 1, 2, 3, 12, 22, 24] and list(kinds) == [0, 0, 1, 2, 0, 0, 0] and list(tx) == [3] * 7 and list(lsn) == [7] * 7 and list(ct) == [5_000_000] * 7
    assert [cell(batch, 0, r) for r in range(7)] == [1, 2, 3, 4, 1, 1, -2147483648]
    assert [cell(batch, 1, r) for r in range(7)] == [b"a", None, b"x", b"d", b"5", "aé\U0001F600 ÿ".encode(), b"wrap"]
    U, H, S = 48, 49, 50
    assert [(r, c) for r, c, _ in errs] == [(4, U), (5, U), (6, U), (7, U), (8, U), (9, U), (10, U), (11, U), (13, U), (14, U), (15, U), (16, U), (17, U), (18, U), (19, S), (20, H), (21, H), (23, U), (25, U)]


# ----------------------------------------------------------------------------------------------------------- GPU parity
def _shape_msgs():
    schema_text = '{"type":"struct","fields":[{"type":"struct","fields":[{"type":"int32","optional":false,"field":"id"},{"type":"string","optional":true,"field":"s"}],"optional":true,"field":"before"},{"type":"struct","fields":[{"type":"int32","optional":false,"field":"id"},{"type":"string","optional":true,"field":"s"}],"optional":true,"field":"after"}]}'
    src = '"source":{"lsn":7,"ts_ms":5,"txId":3,"schema":"public","table":"t","snapshot":"false"}'
    def m(payload, schema=schema_text): return ('{"schema":%s,"payload":%s}' % (schema, payload)).encode()
    msgs = [m('{"before":null,"after":{"id":1,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":2,"s":null},%s,"op":"r"}' % src), m('{"after":{"id":3,"s":"x"},%s,"op":"u"}' % src),
            m('{"before":{"id":4,"s":"d"},"after":null,%s,"op":"d"}' % src), b"", b"{}", b'{"schema":1}', m("null"), m('{"op":"c"}'), m('{"after":{"id":1},%s,"op":"c"}' % src),
            m('{"after":{"id":"1","s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1.5,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":5},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":true},%s,"op":"c"}' % src),
            m('{"after":{"id":1,"s":"a"},%s,"op":"x"}' % src), m('{"after":{"id":1,"s":"a"},%s,"op":7}' % src), m('{"after":{"id":1,"s":"a"},"source":{"lsn":-1},"op":"c"}'), m('{"after":{"id":1,"s":"a"},"source":{"txId":4294967296},"op":"c"}'),
            m('{"after":{"id":1,"s":"a"},%s,"op":"c"}' % src) + b" x", m('{"after":{"id":1,"s":"a"},%s,"op":"c"}' % src, '{"type":"struct","fields":[]}'), m('{"after":{"id":1,"s":"__debezium_unavailable_value"},%s,"op":"c"}' % src),
            m('{"AFTER":{"id":1,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":1,"s":"a\\u00e9\\ud83d\\ude00 \\ud800 \\udc00x"},%s,"op":"c","extra":[1,{"a":null}]}' % src), m('{"after":{"id":9,"s":"dup"},"after":null,%s,"op":"c"}' % src),
            m('{"after":{"id":2147483648,"s":"wrap"},%s,"op":"c"}' % src), b'{"schema":%s,"payload":{"after":{"id":1,"s":"ctl\x01"},"op":"c"}}' % schema_text.encode(),
            b'{"schema":%s,"payload":{"after":{"id":5,"s":"bad utf8 \xff\xc3 \xe2\x82"},"op":"c"}}' % schema_text.encode(), m('{"after":{"id":1,"s":"a"},"after":{"id":2,"s":"b"},%s,"op":"c"}' % src),
            m('{"after":{"id":1,"s":"a"},%s,"op":"c","Op":"u"}' % src), m('{"after":{"id":1,"s":"a"},"source":{"LSN":1},"op":"c"}'), m('{"after":{"id":1,"s":"a","id":77},"source":null,"op":"c","ts_ms":1.5}'),
            m('{"after":{"id":1,"s":"a","id":78},"source":null,"op":null,"op":"c","ts_ms":null}'), b' \n{"schema":%s , "payload" : {"after":{"s":"ws","id": 6 },"op":"c"} }\t' % schema_text.encode(),
            m('{"after":{"id":1e2,"s":"a"},%s,"op":"c"}' % src), m('{"after":{"id":-0,"s":"\\"q\\\\\\/\\b\\f\\n\\r\\t"},%s,"op":"c"}' % src), m('[1]'), b'[{"schema":1}]', m('{"after":[],%s,"op":"c"}' % src)]
    return schema_text, msgs


def _dbz_cmp(eng, po, schema_text, msgs, **kw):
    from transferia_b200 import engine
    from test_gpu_parity import assert_batches_equal
    data = b"".join(msgs); ends = np.cumsum([len(x) for x in msgs]).tolist() if msgs else []
    schema = engine.debezium_table_schema(schema_text)
    table = kw.pop("table", ("public", "t")); allow = kw.pop("_allow_host", 0)
    pid = eng.plan(table[0], table[1], schema, [])
    got, gerr, meta = eng.parse_debezium(pid, data, ends, schema_text, **kw)
    ref, kinds, tx, lsn, ct, rm, rerr, rschema = po.debezium_parse(data, ends, schema_text, use_sr=kw.get("schema_registry", False), schema_id=kw.get("schema_id", 0),
                                                                   table=table if kw.get("check_table") else None)
    assert [{k: c[k] for k in ("name", "type", "key")} for c in rschema] == schema
    extra = [e for e in gerr if e not in rerr]
    if extra and allow:
        # messages the device hands to the host parser although the oracle decides them (magnitudes over 256 bits, float text
        # whose rounding Eisel-Lemire leaves open): bounded in number, then replaced by `{}` so that everything else is compared
        assert all(c == abi.TF_ROWERR_DBZ_HOST for _, c, _ in extra) and len(extra) <= allow, extra
        msgs = list(msgs)
        for r, _, _ in extra: msgs[r] = b"{}"
        return _dbz_cmp(eng, po, schema_text, msgs, table=table, **kw)
    assert gerr == rerr, (gerr, rerr)
    assert list(meta["selection"]) == list(rm)
    assert_batches_equal(got, ref)
    sel = meta["selection"]
    assert list(meta["kinds"][sel]) == list(kinds) and list(meta["tx_id"][sel]) == list(tx) and list(meta["lsn"][sel]) == list(lsn) and list(meta["commit_time"][sel]) == list(ct)
    return got, gerr


@pytest.mark.gpu
def test_device_debezium_canon_and_shapes(eng, po):
    msg = G["messages"][0].encode(); raw = msg.decode(); a = raw.index('"schema":') + len('"schema":'); b = raw.index(',"payload":')
    _dbz_cmp(eng, po, raw[a:b], [msg, msg, b"{}", msg], table=("public", "basic_types"))
    _dbz_cmp(eng, po, raw[a:b], [msg], table=("public", "basic_types"), check_table=True)
    _dbz_cmp(eng, po, raw[a:b], [msg], table=("public", "other"), check_table=True)
    schema_text, msgs = _shape_msgs()
    _dbz_cmp(eng, po, schema_text, msgs)
    _dbz_cmp(eng, po, schema_text, [])
    # schema-registry frames: 0x00 | u32be id | payload
    pay = b'{"after":{"id":1,"s":"sr"},"source":{"lsn":9,"ts_ms":1,"txId":2},"op":"c"}'
    fr = lambda i, p: b"\x00" + int(i).to_bytes(4, "big") + p
    _dbz_cmp(eng, po, schema_text, [fr(7, pay), fr(8, pay), b"\x01" + pay, fr(7, pay)[:4], fr(7, pay) + fr(7, pay), fr(7, pay + b" trailing"), fr(7, b"  " + pay), fr(7, b"{"), b""], schema_registry=True, schema_id=7)


@pytest.mark.gpu
def test_device_debezium_numeric_receivers(eng, po):
    """Decimal / VariableScaleDecimal / Bits / Point over many magnitudes and scales."""
    st = ('{"type":"struct","fields":[{"type":"struct","fields":[%s],"optional":true,"field":"before"},{"type":"struct","fields":[%s],"optional":true,"field":"after"}]}')
    fl = ('{"type":"int64","optional":false,"field":"k"},{"type":"bytes","optional":true,"name":"org.apache.kafka.connect.data.Decimal","version":1,"parameters":{"scale":"3"},"field":"d3"},'
          '{"type":"bytes","optional":true,"name":"org.apache.kafka.connect.data.Decimal","version":1,"field":"d0"},'
          '{"type":"struct","fields":[],"optional":true,"name":"io.debezium.data.VariableScaleDecimal","version":1,"field":"v"},{"type":"bytes","optional":true,"field":"raw"},'
          '{"type":"struct","fields":[],"optional":true,"name":"io.debezium.data.geometry.Point","version":1,"field":"pt"},{"type":"float","optional":true,"field":"f"},{"type":"boolean","optional":true,"field":"b"},{"type":"int8","optional":true,"field":"i8"}')
    schema_text = st % (fl, fl)
    rng = np.random.default_rng(8); msgs = []
    for k in range(400):
        nb = int(rng.integers(0, 36)); mag = bytes(rng.integers(0, 256, nb, dtype=np.uint8)) if k % 7 else b"\x00" * nb
        b64 = base64.b64encode(mag).decode()
        if k % 11 == 0: b64 = b64[:-1] + "!"
        scale = int(rng.integers(-2, 12))
        pt = rng.choice(['{"x":1.5,"y":-2}', '{"x":"s","y":null,"srid":4}', '{"y":1}', '{"x":true,"y":false,"x":7}', '{"x":[1],"y":2}', 'null', '"str"'])
        f = rng.choice(["1.5", "1e400", "0.1", "123456789012345678901234567890", "-0.0", "null", '"1.5"', "2.2250738585072014e-308"])
        msgs.append(('{"schema":%s,"payload":{"after":{"k":%d,"d3":"%s","d0":"%s","v":{"scale":%d,"value":"%s"},"raw":"%s","pt":%s,"f":%s,"b":%s,"i8":%d},"op":"c"}}'
                     % (schema_text, k, b64, b64, scale, b64, b64, pt, f, rng.choice(["true", "false", "null"]), int(rng.integers(-300, 300)))).encode())
    _dbz_cmp(eng, po, schema_text, msgs, _allow_host=80)


def _dbz_fuzz_msgs(n, seed):
    rng = np.random.default_rng(seed)
    schema_text, base = _shape_msgs()
    good = [m for m in base if m.startswith(b'{"schema":') and b'"op":"' in m][:6]
    vals = [b"1", b"-5", b"2147483647", b"2147483648", b"1.0", b"1e3", b'"7"', b"null", b"true", b'"x"', b'"a\\u00e9"', b'"\\ud83d"', b"[1]", b"{}", b'"__debezium_unavailable_value"']
    ops = [b'"c"', b'"r"', b'"u"', b'"d"', b'"x"', b"1", b"null"]
    out = []
    for _ in range(n):
        if rng.random() < 0.5:
            m = bytearray(good[rng.integers(0, len(good))])
            for _ in range(rng.integers(0, 3)):
                p = int(rng.integers(0, len(m))); k = rng.integers(0, 3)
                if k == 0: del m[p]
                elif k == 1: m[p:p] = m[p:p + 1]
                else: m[p] = int(rng.integers(32, 127))
            out.append(bytes(m))
        else:
            after = b'{"id":%s,"s":%s}' % (vals[rng.integers(0, len(vals))], vals[rng.integers(0, len(vals))])
            src = b'{"lsn":%s,"ts_ms":%s,"txId":%s,"schema":"public","table":"t"}' % (vals[rng.integers(0, 4)], vals[rng.integers(0, 8)], vals[rng.integers(0, 6)])
            parts = [b'"after":' + after, b'"before":' + (after if rng.random() < 0.5 else b"null"), b'"source":' + src, b'"op":' + ops[rng.integers(0, len(ops))], b'"ts_ms":' + vals[rng.integers(0, 8)]]
            order = rng.permutation(len(parts))[: rng.integers(2, len(parts) + 1)]
            out.append(b'{"schema":' + schema_text.encode() + b',"payload":{' + b",".join(parts[i] for i in order) + b"}}")
    return schema_text, out


def test_oracle_debezium_fuzz_smoke(po):
    schema_text, msgs = _dbz_fuzz_msgs(2000, 3)
    data = b"".join(msgs); ends = np.cumsum([len(x) for x in msgs]).tolist()
    batch, kinds, tx, lsn, ct, rm, errs, _ = po.debezium_parse(data, ends, schema_text)
    assert batch.nrows + len(errs) == len(msgs) and batch.nrows > 100


@pytest.mark.gpu
def test_device_debezium_fuzz(eng, po):
    """Differential fuzz: mutated envelopes and random payload shapes; the device agrees with the oracle on every message."""
    for seed in (1, 2):
        schema_text, msgs = _dbz_fuzz_msgs(8000, seed)
        _dbz_cmp(eng, po, schema_text, msgs, _allow_host=40)


@pytest.mark.gpu
def test_device_debezium_filter_cast_native_fused(eng, po):
    """BASELINE configs[3] (Debezium CDC -> SQL-predicate filter -> typesystem cast), fused on the device and taken to the ClickHouse
    native block (+ LZ4 frames): schema-registry framed envelopes with a 12-field payload -> tfgpu_parse_debezium with a plan
    (parser.go:34-137 -> filter_rows.go:99-130 -> restore.go / columntypes) vs oracle(parse) -> oracle(push_encode)."""
    from transferia_b200 import engine, workload
    data, ends, schema_text, table = workload.make_debezium_messages(30_000)
    schema = engine.debezium_table_schema(schema_text)
    assert len(schema) == 12
    trs = workload.debezium_transformers()
    pid = eng.plan(table[0], table[1], schema, trs, {"type": "clickhouse"})
    kw = dict(schema_registry=True, schema_id=7)
    ref, kinds, tx, lsn, ct, rm, rerr, _ = po.debezium_parse(data, ends.tolist(), schema_text, use_sr=True, schema_id=7)
    assert rerr == [] and ref.nrows == len(ends) and 0 < int((kinds != 0).sum()) < 0.05 * len(ends)
    refk = abi.Batch(ref.nrows, ref.columns, np.asarray(kinds, dtype=np.uint8))
    want = po.push_encode(refk, po.build_plan(table[0], table[1], schema, trs), abi.TF_WIRE_CH_NATIVE)
    assert 0 < want.rows_out < ref.nrows and len(want.errors) == int((kinds != 0).sum())          # update / delete: "Found non-supported kind"
    res, meta = eng.parse_debezium(pid, data, ends, schema_text, wire_fmt=abi.TF_WIRE_CH_NATIVE, **kw)
    assert res.rows_in == len(ends) and res.rows_out == want.rows_out and res.errors == want.errors
    assert res.wire == want.raw
    assert list(meta["kinds"]) == list(kinds) and list(meta["lsn"]) == list(lsn)
    res_lz, _ = eng.parse_debezium(pid, data, ends, schema_text, wire_fmt=abi.TF_WIRE_CH_NATIVE_LZ4, **kw)
    raw, nf = po.ch_decode_frames(res_lz.wire)
    assert raw == want.raw and nf == res_lz.n_frames

