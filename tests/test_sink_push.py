"""tfgpu_sink_push — Sinker.Push through the always-on middleware: sequencing and counters against
oracle/middleware_oracle.py (a line-by-line restatement of transformation.Push / NonRowSeparator / Filter / Statistician) on CPU, and the
whole path (rows -> transpose -> device chain -> frames -> ClickHouse native writer -> server peer) on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ch_peer
from oracle import middleware_oracle as mo
from transferia_b200 import abi, engine, rows, sink, workload
from transferia_b200.rows import ChangeItem, go

K = rows
SCHEMA_A = [{"name": "id", "type": "int32", "key": True}, {"name": "v", "type": "utf8"}]
SCHEMA_B = [{"name": "x", "type": "int64"}]
TABLES = [("public", "a", SCHEMA_A), ("public", "b", SCHEMA_B), ("public", "__consumer_keeper", SCHEMA_B), ("public", "a", SCHEMA_A + [{"name": "w", "type": "int8"}])]


def _items(spec):
    """spec: [(kind, table)] -> ChangeItems with values fitting the table's schema."""
    out = []
    for i, (kind, t) in enumerate(spec):
        vals = None
        if kind <= 2:
            vals = {0: [go.int32(i), go.string(f"v{i}")], 1: [go.int64(i)], 2: [go.int64(i)], 3: [go.int32(i), go.string("z"), go.int8(1)]}[t]
        out.append(ChangeItem(kind, t, vals, id=i, lsn=1000 + i, commit_time=0 if i % 7 == 3 else 5_000 + 13 * ((i * 7) % 11), size_read=1, size_values=10 + i))
    return out


def _oracle_items(items):
    return [{"kind": it.kind, "table": it.table, "index": i, "commit_time": it.commit_time, "size": it.size_values} for i, it in enumerate(items)]


def _zero_stats():
    return {k: 0 for k in ("pushes", "downstream_pushes", "change_items_pushed", "row_events_pushed", "inflight_bytes", "filter_dropped", "transform_dropped",
                           "transform_errors", "max_commit_time", "min_commit_time", "without_commit_time", "wire_bytes")}


def _events(s):
    names = {sink.EV_ROWS: "rows", sink.EV_ITEM: "item", sink.EV_ERRORS: "errors"}
    return [{"type": names[e["type"]], "table": e["table"], "items": e["items"]} for e in s.events]


def test_nonrow_separator_reference_cases():
    """pkg/middlewares/nonrow_separator_test.go: every item reaches the sink; runs of rows stay together, every other item travels alone."""
    s = sink.Sink()
    cases = [
        ([K.KIND_INSERT] * 3 + [K.KIND_UPDATE, K.KIND_DELETE, K.KIND_INSERT], [6]),                                        # SingleBatchOnlyRows
        ([K.KIND_INSERT, K.KIND_INSERT, K.KIND_UPDATE, K.KIND_DROP_TABLE, K.KIND_DELETE, K.KIND_INSERT], [3, 1, 2]),         # SingleBatchWithNonRows
        ([K.KIND_DROP_TABLE], [1]), ([K.KIND_INSERT, K.KIND_DROP_TABLE], [1, 1]), ([], []),                                # TwoBatchesWith / WithoutRows
        ([K.KIND_DDL] * 4, [1, 1, 1, 1]),                                                                                  # MultipleNonRows
    ]
    for kinds, want in cases:
        s.events.clear()
        its = _items([(k, 0) for k in kinds])
        s.push(rows.RowsImage(its, TABLES))
        assert [e["n_items"] for e in s.events] == want, kinds
        assert sum(e["n_items"] for e in s.events) == len(kinds)
        assert [x for e in s.events for x in e["items"]] == list(range(len(kinds)))
    s.close()


def test_push_sequence_and_counters_equal_the_oracle():
    rng = np.random.default_rng(5)
    kinds = [K.KIND_INSERT] * 6 + [K.KIND_UPDATE, K.KIND_DELETE, K.KIND_INIT_TABLE_LOAD, K.KIND_DONE_TABLE_LOAD, K.KIND_TRUNCATE, K.KIND_DDL, K.KIND_SYNCHRONIZE, K.KIND_OTHER]
    s = sink.Sink(system_tables=["__consumer_keeper", "__tm_keeper"])
    want_stats = _zero_stats()
    for rnd in range(30):
        spec = [(int(rng.choice(kinds)), int(rng.choice([0, 0, 0, 1, 2, 3]))) for _ in range(int(rng.integers(0, 60)))]
        its = _items(spec)
        s.events.clear()
        s.push(rows.RowsImage(its, TABLES))
        want = mo.sink_push(_oracle_items(its), [(t[0], t[1]) for t in TABLES], [], {}, ["__consumer_keeper", "__tm_keeper"],
                            lambda table, run: (run, []), True, want_stats)
        assert _events(s) == want, (rnd, spec)
        got = s.stats(); got.pop("wire_bytes"); w = dict(want_stats); w.pop("wire_bytes")
        assert got.pop("metering_output_rows") == w["change_items_pushed"] and got.pop("metering_input_rows") >= w["change_items_pushed"]   # metering.go:42-48,63-69
        assert got == w, rnd
        # row runs arrive columnar, in item order
        for e in s.events:
            if e["type"] == sink.EV_ROWS:
                assert list(e["columns"][0]) == e["items"] and e["out"][1] in ("a", "b")
    assert s.stats()["filter_dropped"] > 0 and s.stats()["without_commit_time"] > 0
    s.close()


def test_filter_can_be_switched_off_and_downstream_errors_propagate():
    its = _items([(K.KIND_INSERT, 2), (K.KIND_INSERT, 0), (K.KIND_DDL, 2)])
    s = sink.Sink(system_tables=["__consumer_keeper"], exclude_system_tables=False)      # destinations that rely on system tables (sink_factory.go:87-89)
    s.push(rows.RowsImage(its, TABLES))
    assert [e["n_items"] for e in s.events] == [1, 1, 1] and s.stats()["filter_dropped"] == 0
    s.close()
    calls = []
    s = sink.Sink(downstream=lambda ev: (calls.append(ev["type"]), 7 if len(calls) == 2 else 0)[1])
    with pytest.raises(engine.EngineError) as ei:
        s.push(rows.RowsImage(_items([(K.KIND_INSERT, 0), (K.KIND_DDL, 0), (K.KIND_INSERT, 0)]), TABLES))
    assert ei.value.rc == 7 and len(calls) == 2
    st = s.stats()
    assert st["downstream_pushes"] == 1 and st["change_items_pushed"] == 1                # the failed Push is not counted (statistician.go:60)
    s.close()


def test_transformers_need_the_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError) as ei:
        sink.Sink(transformers=[{"filter_rows": {"filter": "id > 1"}}])
    assert ei.value.rc == engine.TF_E_FATAL_NODEVICE
    with pytest.raises(engine.EngineError):
        sink.Sink(wire_fmt=abi.TF_WIRE_CH_NATIVE_LZ4)


@pytest.mark.gpu
def test_push_with_transformers_on_the_device_equals_the_oracle(eng, po):
    """skip_events / rename_tables / filter_rows over mixed kinds and two tables; the row runs come back as columnar batches (wire_fmt 0)
    that equal the oracle's push_columns over the same rows; sequencing and counters equal the middleware oracle."""
    # filter_rows before rename_tables: its Apply re-checks the table filter on the item's CURRENT id (a renamed item would pass unfiltered)
    trs = [{"skip_events": {"tables": {"includeTables": ["^public.a$"]}, "events": ["truncate", "delete"]}},
           {"filter_rows": {"tables": {"includeTables": ["^public.a$"]}, "filter": "id > 10"}},
           {"rename_tables": {"renameTables": [{"originalName": {"nameSpace": "public", "name": "a"}, "newName": {"nameSpace": "dst", "name": "a2"}}]}}]
    spec = [(K.KIND_INIT_TABLE_LOAD, 0)] + [(K.KIND_INSERT, 0)] * 9 + [(K.KIND_TRUNCATE, 0)] + [(K.KIND_INSERT, 0)] * 12 + [(K.KIND_INSERT, 1)] * 5 + \
           [(K.KIND_DELETE, 0), (K.KIND_DDL, 0), (K.KIND_INSERT, 0), (K.KIND_DONE_TABLE_LOAD, 0), (K.KIND_INSERT, 2)]
    its = _items(spec)
    s = sink.Sink(eng, transformers=trs, system_tables=["__consumer_keeper"])
    s.push(rows.RowsImage(its, TABLES))

    def apply(table, run):
        if table != 0:
            return run, []
        kept = [x for x in run if x["kind"] != K.KIND_DELETE and x["index"] > 10]      # skip_events drops deletes, id == item index
        return kept, []
    want_stats = _zero_stats()
    suit_a = lambda ns, name: (ns, name) == ("public", "a")
    want = mo.sink_push(_oracle_items(its), [(t[0], t[1]) for t in TABLES], [(suit_a, ["truncate", "delete"])], {("public", "a"): ("dst", "a2")},
                        ["__consumer_keeper"], apply, True, want_stats)
    got = _events(s)
    for g, w in zip(got, want):                                                             # rows dropped by the device: the event carries the count, not the indexes
        assert g["type"] == w["type"] and g["table"] == w["table"]
        assert g["items"] == w["items"] or (g["items"] is None and s.events[got.index(g)]["n_items"] == len(w["items"]))
    assert len(got) == len(want)
    st = s.stats(); st.pop("wire_bytes"); want_stats.pop("wire_bytes"); st.pop("metering_input_rows"); st.pop("metering_output_rows")
    for k in ("inflight_bytes", "max_commit_time", "min_commit_time", "without_commit_time"):   # per-item figures need the indexes the device does not return
        st.pop(k); want_stats.pop(k)
    assert st == want_stats
    ev_rows = [e for e in s.events if e["type"] == sink.EV_ROWS and e["table"] == 0]
    assert all(e["out"] == ("dst", "a2") for e in ev_rows) and [e["out"] for e in s.events if e["type"] == sink.EV_ITEM][0] == ("dst", "a2")
    assert sorted(int(v) for e in ev_rows for v in e["columns"][0]) == [i for i, (k, t) in enumerate(spec) if t == 0 and k == K.KIND_INSERT and i > 10]
    s.close()


@pytest.mark.gpu
def test_push_rows_to_clickhouse_end_to_end(eng, po):
    """Sinker.Push of ClickBench-shaped rows: transpose -> filter_rows + cast + native block + LZ4 frames on the device -> INSERT over the
    native protocol; the server peer's copy decodes to the oracle's block and the statement is the one doOperation builds."""
    batch, schema = workload.make_hits_batch(5000, seed=12)
    trs = workload.headline_transformers(workload.counterid_threshold(batch, schema))
    items = [ChangeItem(K.KIND_INIT_TABLE_LOAD, 0)] + rows.items_from_batch(batch) + [ChangeItem(K.KIND_DONE_TABLE_LOAD, 0)]
    ref = po.push_encode(batch, po.build_plan("public", "hits", schema, trs), abi.TF_WIRE_CH_NATIVE, eng.frame_bytes)
    names = [c["name"] for c in schema]
    cli, srv = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    peer = ch_peer.Peer(srv, [(n, "String") for n in names], expect_raw_len=[len(ref.raw)]); peer.start()
    w = sink.ClickHouseWriter(cli, database="analytics", read_timeout_ms=60000)
    s = sink.Sink(eng, transformers=trs, wire_fmt=abi.TF_WIRE_CH_NATIVE_LZ4, database="analytics", clickhouse=w)
    s.push(rows.RowsImage(items, [("public", "hits", schema)]))
    peer.join(60)
    assert peer.error is None, peer.error
    assert peer.queries[0]["body"] == sink.insert_query("analytics", "hits", names)
    raw, _ = po.ch_decode_frames(peer.blocks[0][0])
    assert raw == ref.raw
    assert [e["type"] for e in s.events] == [sink.EV_ITEM, sink.EV_ITEM]                    # the control items went to the callback, the rows to the socket
    st = s.stats()
    assert st["row_events_pushed"] == ref.rows_out and st["change_items_pushed"] == ref.rows_out + 2 and st["transform_dropped"] == 5000 - ref.rows_out
    assert st["wire_bytes"] == len(peer.blocks[0][0])
    s.close(); w.close(); cli.close()


@pytest.mark.gpu
def test_push_cdc_rows_to_debezium_messages(eng, po):
    """Sinker.Push with the queue Debezium serializer as destination (wire_fmt TF_WIRE_DEBEZIUM): insert / update / delete items with OldKeys,
    ID / LSN / CommitTime in the row form -> transpose -> tfgpu_emit_debezium_crud; the messages equal the oracle's Emitter.emitKV byte for byte."""
    from test_debezium_emit import OPTS
    rng = np.random.default_rng(9); n = 2000
    schema = [{"name": "id", "type": "int64", "key": True}, {"name": "k2", "type": "utf8", "key": True}, {"name": "name", "type": "utf8"}, {"name": "x", "type": "double"},
              {"name": "ts", "type": "timestamp"}]
    ids = rng.integers(0, 50, n); k2 = [b"k%d" % v for v in rng.integers(0, 5, n)]
    def mk(ids_, k2_, salt):
        return abi.Batch(n, [abi.fixed_to_column(abi.TF_INT64, ids_), abi.strings_to_column(abi.TF_UTF8, k2_),
                             abi.strings_to_column(abi.TF_UTF8, [None if (i + salt) % 7 == 0 else b"n<%d>" % (i * salt) for i in range(n)]),
                             abi.fixed_to_column(abi.TF_DOUBLE, rng.random(n) * 1e3, [(i + salt) % 11 == 0 for i in range(n)]),
                             abi.fixed_to_column(abi.TF_TIMESTAMP, rng.integers(0, 2**31, n), None, rng.integers(0, 10**9, n).astype(np.uint32))])
    kinds = rng.integers(0, 3, n).astype(np.uint8)
    b = mk(ids, k2, 1); b.kinds = kinds
    same = rng.random(n) < 0.6
    old = mk(np.where(same, ids, ids + 1), [a if s_ else a + b"x" for a, s_ in zip(k2, rng.random(n) < 0.8)], 3)
    has = ((rng.random(n) < 0.9) & (kinds != 0)).astype(np.uint8)           # inserts carry no OldKeys
    present = [0, 1, 2]
    meta = {"id": rng.integers(0, 2**32, n, dtype=np.uint32), "lsn": rng.integers(0, 2**62, n, dtype=np.uint64), "commit_time": rng.integers(1, 2**62, n, dtype=np.uint64)}
    items = rows.items_from_batch(b); olds = rows.items_from_batch(old)
    for r, it in enumerate(items):
        it.id, it.lsn, it.commit_time = int(meta["id"][r]), int(meta["lsn"][r]), int(meta["commit_time"][r])
        if has[r]:
            it.old_keys = {c: olds[r].values[c] for c in present}
    opts = {k: v for k, v in OPTS.items()}
    plan = po.build_plan("public", "t", schema, [])
    want = po.debezium_emit(b, plan, opts, meta, old=old, old_present=present, old_row_has=has, want_msg_sizes=True)
    s = sink.Sink(eng, wire_fmt=abi.TF_WIRE_DEBEZIUM, debezium=opts)
    s.push(rows.RowsImage(items, [("public", "t", schema)]))
    ev = [e for e in s.events if e["type"] == sink.EV_ROWS]
    assert len(ev) == 1 and ev[0]["n_items"] == n and ev[0]["wire"] == want[0]
    assert np.array_equal(ev[0]["msg_sizes"][:, 0], want[4][:, 0]) and set(want[4][:, 0].tolist()) >= {1, 2, 3}
    st = s.stats()
    assert st["row_events_pushed"] == n and st["wire_bytes"] == len(want[0])
    s.close()


def _split_sink(columns, splitter, tables=None, **kw):
    cfg = {"columns": columns, "splitter": splitter}
    if tables:
        cfg["tables"] = tables
    return sink.Sink(transformers=[{"table_splitter": cfg}], **kw)


def test_table_splitter_reference_cases():
    """registry/table_splitter/table_splitter_test.go:68-118: the generated table names, through Sinker.Push (host-level transformers need no
    device: skip_events / rename_tables / table_splitter act on kinds and table names only)."""
    s1 = [{"name": "column1", "type": "string", "key": True}, {"name": "column2", "type": "int64"}, {"name": "column3", "type": "int32"}, {"name": "column4", "type": "boolean"}]
    s2 = [{"name": "column1", "type": "string"}, {"name": "column2", "type": "date"}, {"name": "column3", "type": "double"}, {"name": "column4", "type": "float"}]
    f1 = lambda a, b, c, d: [go.string(a), go.int64(b), go.int32(c), go.bool(d)]
    cases = [([], "_", "table1", s1, f1("hello", 123, 321, False), "table1"),
             (["column1"], "_", "table2", s1, f1("hello", 456, 654, False), "table2_hello"),
             (["column1"], "/", "table3", s1, f1("hello", 789, 987, False), "table3/hello"),
             (["column1", "column2"], "$", "table4", s1, f1("hello", 345, 543, False), "table4$hello$345"),
             (["column2", "column1"], "+", "table5", s1, f1("hello", 678, 876, False), "table5+678+hello"),
             (["column1", "column2"], "", "table6", s1, f1("helloworld", 234, 432, False), "table6/helloworld/234"),
             # 2023-08-31T16:59:07+03:00 -> the UTC date; float32 2.71828 prints with its own shortest digits
             (["column4", "column2"], "__", "table7", s2, [go.string("helloworld"), go.time(1693490347), go.float64(3.14), go.float32(2.71828)], "table7__2.71828__2023-08-31")]
    for cols, sp, tname, schema, vals, want in cases:
        s = _split_sink(cols, sp)
        s.push(rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, vals)], [("db", tname, schema)]))
        assert [e["out"][1] for e in s.events] == [want], (cols, sp)
        assert mo.generate_table_name(tname, cols, sp, dict(zip([c["name"] for c in schema], vals)), {c["name"]: c["type"] for c in schema}) == want
        s.close()
    # Suitable: the table filter (table_splitter_test.go:121-169) — a table outside it keeps its name
    s = _split_sink(["column1"], "_", tables={"includeTables": ["^db\\.table3$"]})
    s.push(rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, f1("x", 1, 2, True)), ChangeItem(K.KIND_INSERT, 1, f1("y", 1, 2, True))], [("db", "table1", s1), ("db", "table3", s1)]))
    assert [e["out"][1] for e in s.events] == ["table1", "table3_y"]
    s.close()


def test_table_splitter_groups_rows_and_renames_control_items():
    rng = np.random.default_rng(11)
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "region", "type": "utf8"}, {"name": "day", "type": "date"}, {"name": "w", "type": "double"}]
    regions = [b"eu", b"us", b"apac"]; days = [19000 * 86400, 19001 * 86400 + 3600]
    items = [ChangeItem(K.KIND_INIT_TABLE_LOAD, 0)]
    for i in range(400):
        v = [go.int32(i), go.nil if i % 37 == 0 else go.string(regions[int(rng.integers(0, 3))]), go.time(days[int(rng.integers(0, 2))], 5), go.float64(float(rng.integers(0, 3)) * 0.25 + 1e6 * (i % 2))]
        items.append(ChangeItem(K.KIND_UPDATE if i % 5 == 0 else K.KIND_INSERT, 0, v))
    items.append(ChangeItem(K.KIND_DONE_TABLE_LOAD, 0))
    trs = [{"skip_events": {"events": ["update"]}},
           {"rename_tables": {"renameTables": [{"originalName": {"nameSpace": "public", "name": "t"}, "newName": {"nameSpace": "dst", "name": "t2"}}]}},
           {"table_splitter": {"columns": ["region", "day", "w", "nope"], "splitter": "-"}}]
    s = sink.Sink(transformers=trs)
    s.push(rows.RowsImage(items, [("public", "t", schema)]))
    types = {c["name"]: c["type"] for c in schema}; names = [c["name"] for c in schema]
    want_groups = {}
    for k, it in enumerate(items):
        if it.kind == K.KIND_INSERT:
            want_groups.setdefault(mo.generate_table_name("t2", ["region", "day", "w", "nope"], "-", dict(zip(names, it.values)), types), []).append(k)
    ev_rows = [e for e in s.events if e["type"] == sink.EV_ROWS]
    assert {e["out"][1]: e["items"] for e in ev_rows} == want_groups and len(want_groups) >= 12           # order of first appearance, rows in item order
    assert [e["out"][1] for e in ev_rows] == list(want_groups)
    assert any("<nil>" in n for n in want_groups) and any("e+06" in n for n in want_groups)
    ctl = [e for e in s.events if e["type"] == sink.EV_ITEM]
    assert [e["out"] for e in ctl] == [("dst", "t2-<nil>-<nil>-<nil>")] * 2 and [e["items"] for e in ctl] == [[0], [len(items) - 1]]
    st = s.stats()
    assert st["transform_dropped"] == 80 and st["row_events_pushed"] == 320 and st["change_items_pushed"] == 322
    for e in ev_rows:                                                                                       # the run arrives columnar: ids of that group
        assert list(e["columns"][0]) == [items[k].values[0][1] for k in e["items"]]
    s.close()
    # plain tfgpu_plan callers are refused: one block cannot hold several tables
    with pytest.raises(engine.EngineError) as ei:
        engine.plan_validate("public", "t", schema, [{"table_splitter": {"columns": ["region"]}}])
    assert ei.value.rc == -2 and "tfgpu_sink_push" in str(ei.value)


@pytest.mark.gpu
def test_table_splitter_with_device_chain(eng, po):
    """filter_rows on the device + table_splitter on the host: every generated table gets its own native block, equal to the oracle's block over
    that table's rows; the plan the sink builds marks the splitter step so that the device ignores it."""
    batch, schema = workload.make_hits_batch(3000, seed=23)
    names = [c["name"] for c in schema]
    ci = next(k for k, c in enumerate(batch.columns) if c.type == abi.TF_INT16 and 2 <= len(np.unique(np.asarray(c.values))) <= 6)
    flt = {"filter_rows": {"filter": f"{names[0]} > 0"}}
    s = sink.Sink(eng, transformers=[flt, {"table_splitter": {"columns": [names[ci]], "splitter": "_"}}], wire_fmt=abi.TF_WIRE_CH_NATIVE)
    s.push(rows.RowsImage(rows.items_from_batch(batch), [("public", "hits", schema)]))
    ev = [e for e in s.events if e["type"] == sink.EV_ROWS]
    vals = np.asarray(batch.columns[ci].values)
    assert [e["out"][1] for e in ev] == [f"hits_{v}" for v in dict.fromkeys(vals.tolist())]
    plan = po.build_plan("public", "hits", schema, [flt]); pool = rows.Columnar()
    for e in ev:
        v = int(e["out"][1].split("_")[1])
        part, _ = pool.gather(batch, (vals == v).astype(np.uint8))
        want = po.push_encode(part, plan, abi.TF_WIRE_CH_NATIVE)
        assert e["wire"] == want.raw and e["n_items"] == want.rows_out
    assert s.stats()["row_events_pushed"] == po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE).rows_out
    s.close(); pool.close()


def test_host_level_transformer_fuzz_against_the_oracle():
    """Random batches through skip_events + rename_tables + table_splitter (no device needed) against the middleware oracle extended with the
    splitter's grouping: the same downstream pushes in the same order, the same counters."""
    rng = np.random.default_rng(21)
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "g", "type": "utf8"}]
    tables = [("public", "a", schema), ("public", "b", schema), ("public", "__consumer_keeper", schema)]
    kinds = [K.KIND_INSERT] * 6 + [K.KIND_UPDATE, K.KIND_DELETE, K.KIND_INIT_TABLE_LOAD, K.KIND_TRUNCATE, K.KIND_DDL]
    trs = [{"skip_events": {"tables": {"includeTables": ["^public.a$"]}, "events": ["truncate", "delete"]}},
           {"rename_tables": {"renameTables": [{"originalName": {"nameSpace": "public", "name": "b"}, "newName": {"nameSpace": "x", "name": "b2"}}]}},
           # (an excludeTables regexp would have to match BOTH name variants, public.t and "public"."t", to exclude a table: MatchAnyTableNameVariant,
           #  transformer_common.go:9-33 — the TODO in table_splitter_test.go:160-161 is that quirk; includeTables is unambiguous)
           {"table_splitter": {"tables": {"includeTables": ["^public.a$", "^public.b$"]}, "columns": ["g"], "splitter": "."}}]
    s = sink.Sink(transformers=trs, system_tables=["__consumer_keeper"])
    types = {"id": "int32", "g": "utf8"}
    for rnd in range(25):
        spec = [(int(rng.choice(kinds)), int(rng.choice([0, 0, 1, 2]))) for _ in range(int(rng.integers(0, 50)))]
        items = []
        for i, (kind, t) in enumerate(spec):
            vals = [go.int32(i), go.nil if i % 9 == 0 else go.string(["p", "q", "r"][int(rng.integers(0, 3))])] if kind <= 2 else None
            items.append(ChangeItem(kind, t, vals, commit_time=100 + i, size_values=1))
        s.events.clear()
        s.push(rows.RowsImage(items, tables))
        # expected: tables in order of first appearance; per table the item sequence with delete / truncate of `a` gone; maximal row runs; every
        # run split by generated name in order of first appearance; control items renamed (and split: "<nil>" per known column); system table dropped
        want = []
        order = list(dict.fromkeys(t for _, t in spec))
        for t in order:
            ns, name, _ = tables[t]
            out_name = "b2" if name == "b" else name
            seq = []
            for i, (kind, tt) in enumerate(spec):
                if tt != t: continue
                if name == "a" and kind in (K.KIND_DELETE, K.KIND_TRUNCATE): continue
                if kind <= 2:
                    if seq and seq[-1][0] == "run": seq[-1][1].append(i)
                    else: seq.append(("run", [i]))
                else:
                    seq.append(("item", i))
            for ent in seq:
                if name == "__consumer_keeper":
                    continue                                                   # not split (excluded), then dropped by the system-table filter
                if ent[0] == "item":
                    want.append(("item", out_name + ".<nil>", [ent[1]]))
                else:
                    groups = {}
                    for i in ent[1]:
                        groups.setdefault(mo.generate_table_name(out_name, ["g"], ".", {"id": items[i].values[0], "g": items[i].values[1]}, types), []).append(i)
                    want += [("rows", nm, idx) for nm, idx in groups.items()]
        got = [({sink.EV_ROWS: "rows", sink.EV_ITEM: "item"}[e["type"]], e["out"][1], e["items"]) for e in s.events]
        assert got == want, (rnd, spec)
    st = s.stats()
    assert st["filter_dropped"] > 0 and st["transform_dropped"] > 0 and st["change_items_pushed"] == st["metering_output_rows"]
    s.close()


def test_batch_stats_reference_case():
    """pkg/stats/sink_wrapper_util_test.go:18-28 (TestBatchStats, default case): two inserts with CommitTime 19 and 99 and Size.Values 1 and 2 ->
    `oldestTime` 99 (sic: the latest), `freshestTime` 19, 2 row events, 3 bytes. A Synchronize item counts for the times but not as a row event."""
    s = sink.Sink()
    s.push(rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, [go.int32(1), go.string("a")], commit_time=19, size_read=1, size_values=1),
                           ChangeItem(K.KIND_INSERT, 0, [go.int32(2), go.string("b")], commit_time=99, size_read=1, size_values=2)], TABLES))
    st = s.stats()
    assert (st["max_commit_time"], st["min_commit_time"], st["row_events_pushed"], st["inflight_bytes"]) == (99, 19, 2, 3)
    s.push(rows.RowsImage([ChangeItem(K.KIND_SYNCHRONIZE, 0, None, commit_time=0)], TABLES))
    st = s.stats()
    assert st["row_events_pushed"] == 2 and st["without_commit_time"] == 1 and st["change_items_pushed"] == 3 and st["inflight_bytes"] == 3
    s.close()


def test_control_items_without_a_table_schema():
    """DDL / drop_table items may carry no TableSchema: they travel alone like every non-row item, renamed by rename_tables; table_splitter
    has no columns to read there and leaves the name alone."""
    trs = [{"rename_tables": {"renameTables": [{"originalName": {"nameSpace": "public", "name": "ghost"}, "newName": {"nameSpace": "x", "name": "ghost2"}}]}},
           {"table_splitter": {"columns": ["id"], "splitter": "_"}}]
    for t in (None, trs):
        s = sink.Sink(transformers=t)
        s.push(rows.RowsImage([ChangeItem(K.KIND_DDL, 1), ChangeItem(K.KIND_INSERT, 0, [go.int32(1), go.string("a")]), ChangeItem(K.KIND_DROP_TABLE, 1)],
                              [("public", "a", SCHEMA_A), ("public", "ghost", None)]))
        # tables in order of first appearance: `ghost` (item 0) before `a`
        assert [(e["type"], e["items"]) for e in s.events] == [(sink.EV_ITEM, [0]), (sink.EV_ITEM, [2]), (sink.EV_ROWS, [1])]
        assert [e["out"] for e in s.events] == ([("public", "ghost")] * 2 + [("public", "a")] if t is None else [("x", "ghost2")] * 2 + [("public", "a_1")])
        s.close()


def test_dispatcher_delivers_in_submission_order():
    """tfgpu_dispatcher over three sinks: batches are dealt round-robin and worked on concurrently, but what reaches the destination arrives batch
    by batch in submission order, whichever sink finished first (slow and fast destinations mixed); a failing Push surfaces on its own batch."""
    import threading, time
    log, lock = [], threading.Lock()
    def downstream(k):
        def f(ev):
            time.sleep(0.004 * ((ev["items"][0] * 7 + k) % 3))      # uneven destinations
            with lock:
                log.append((k, ev["type"], tuple(ev["columns"][0]) if ev.get("columns") else None))
            return 9 if (ev.get("columns") and 404 in list(ev["columns"][0])) else 0
        return f
    sinks = [sink.Sink(downstream=downstream(k)) for k in range(3)]
    d = sink.Dispatcher(sinks)
    images, seqs = [], []
    for b in range(14):
        ids = [404] if b == 9 else list(range(b * 10, b * 10 + 3 + b % 4))
        its = [ChangeItem(K.KIND_INSERT, 0, [go.int32(i), go.string("x")]) for i in ids] + ([ChangeItem(K.KIND_DDL, 0)] if b % 5 == 2 else [])
        img = rows.RowsImage(its, TABLES); images.append((ids, b % 5 == 2)); seqs.append(d.submit(img))
    assert seqs == list(range(14))
    failed = []
    for s_ in seqs:
        try:
            d.wait(s_)
        except engine.EngineError as ex:
            failed.append((s_, ex.rc))
    assert failed == [(9, 9)]
    want = []
    for b, (ids, ddl) in enumerate(images):
        want.append((b % 3, sink.EV_ROWS, tuple(ids)))
        if ddl: want.append((b % 3, sink.EV_ITEM, None))
    assert log == want
    assert d.drain() == 9 and d.drain() == 0
    assert sum(s_.stats()["row_events_pushed"] for s_ in sinks) == sum(len(i) for i, _ in images) - 1      # the failed Push is not counted
    d.close()
    for s_ in sinks:
        s_.push(rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, [go.int32(1), go.string("y")])], TABLES))      # usable on their own again
        s_.close()


def test_table_splitter_float_and_time_text_fuzz():
    """The host's SerializeToString of float32 / float64 (fmt %v = strconv 'g' with the shortest digits: exponent form below 1e-4 and from
    1e+06 on), int64 extremes, dates and timestamps with nanoseconds, against the oracle's restatement (numpy's shortest repr + Go's layout
    rules): two implementations that share no code."""
    rng = np.random.default_rng(33)
    f64 = np.concatenate([rng.standard_normal(150) * 10.0 ** rng.integers(-12, 25, 150), [0.0, -0.0, 1e5, 1e6, 999999.0, 123456.7, 1e-4, 1e-5, 0.000123, 1e21, 1e22, 5e-324, 1.7976931348623157e308, 100.0, 2.5]])
    f32 = np.concatenate([(rng.standard_normal(150) * 10.0 ** rng.integers(-8, 20, 150)).astype(np.float32), np.array([2.71828, 16777216.0, 1e6, 999999.0, 1e-5, 0.1, 3.4028235e38], np.float32)])
    secs = np.concatenate([rng.integers(-10**9, 4 * 10**9, 60), [0, -1, 86399, 951782400, 1709164800]])
    schema = [{"name": "d", "type": "double"}, {"name": "f", "type": "float"}, {"name": "i", "type": "int64"}, {"name": "day", "type": "date"}, {"name": "ts", "type": "timestamp"}]
    types = {c["name"]: c["type"] for c in schema}
    items = []
    for k in range(max(len(f64), len(f32))):
        items.append(ChangeItem(K.KIND_INSERT, 0, [go.float64(f64[k % len(f64)]), go.float32(f32[k % len(f32)]), go.int64([-2**63, 2**63 - 1, 0, -7][k % 4]),
                                                   go.time(int(secs[k % len(secs)])), go.time(int(secs[(k * 7) % len(secs)]), [0, 5, 500_000_000, 123_456_789][k % 4])]))
    for cols in (["d"], ["f"], ["i", "day"], ["ts"]):
        s = sink.Sink(transformers=[{"table_splitter": {"columns": cols, "splitter": "|"}}])
        s.push(rows.RowsImage(items, [("", "t", schema)]))
        got = {}
        for e in s.events:
            for i in e["items"]: got[i] = e["out"][1]
        for i, it in enumerate(items):
            want = mo.generate_table_name("t", cols, "|", dict(zip([c["name"] for c in schema], it.values)), types)
            assert got[i] == want, (cols, it.values, got[i], want)
        s.close()


def test_updatable_clickhouse_table_rows():
    """cfg.updateable (model.ChSinkParams.IsUpdateable): inserts get (CommitTime, 0) behind their values, deletes are rebuilt from OldKeys with
    (CommitTime, CommitTime) — buildChangeItemArgs / buildDeleteKindArgs, sink_table.go:397-432 — and go down as ordinary rows of the table
    extended by `__data_transfer_commit_time` / `__data_transfer_delete_time`; updates and transformer chains are refused."""
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "k2", "type": "utf8", "key": True}, {"name": "v", "type": "double"}]
    items = []
    for i in range(40):
        if i % 4 == 3:
            items.append(ChangeItem(K.KIND_DELETE, 0, {}, {0: go.int32(i), 1: go.string(f"k{i}")}, commit_time=1000 + i))
        else:
            items.append(ChangeItem(K.KIND_INSERT, 0, [go.int32(i), go.string(f"k{i}"), go.nil if i % 5 == 0 else go.float64(i / 4)], commit_time=1000 + i))
    s = sink.Sink(updateable=True)
    s.push(rows.RowsImage(items, [("public", "t", schema)]))
    (ev,) = [e for e in s.events if e["type"] == sink.EV_ROWS]
    b = ev["batch"]
    assert ev["items"] == list(range(40)) and b.nrows == 40 and len(b.columns) == 5 and b.kinds is None
    want = [mo.updatable_args(it.kind, [None if v[0] == 0 else v[1] for v in (it.values or [])], {c: v[1] for c, v in (it.old_keys or {}).items()}, it.commit_time, 3) for it in items]
    ids, valid0 = np.asarray(b.columns[0].values), np.ones(40, bool) if b.columns[0].validity is None else np.unpackbits(np.asarray(b.columns[0].validity), bitorder="little")[:40].astype(bool)
    assert [int(ids[r]) if valid0[r] else None for r in range(40)] == [w[0] for w in want]
    ln = np.asarray(b.columns[1].offsets).astype(np.int64); heap = np.asarray(b.columns[1].heap).tobytes(); off = np.concatenate([[0], np.cumsum(ln)])
    assert [heap[off[r]:off[r + 1]] for r in range(40)] == [w[1] for w in want]
    v2 = np.unpackbits(np.asarray(b.columns[2].validity), bitorder="little")[:40].astype(bool); dv = np.asarray(b.columns[2].values)
    assert [float(dv[r]) if v2[r] else None for r in range(40)] == [w[2] for w in want]          # nil for every delete: the column is not in OldKeys
    assert b.columns[3].type == abi.TF_UINT64 and list(np.asarray(b.columns[3].values)) == [w[3] for w in want]
    assert list(np.asarray(b.columns[4].values)) == [w[4] for w in want] and sum(1 for w in want if w[4]) == 10
    s.close()
    # the INSERT statement of such a table lists the two system columns once (sink_table.go:643-648)
    d = engine.plan_validate("public", "t", schema + [{"name": "__data_transfer_commit_time", "type": "uint64", "required": True}, {"name": "__data_transfer_delete_time", "type": "uint64", "required": True}], [], {"type": "clickhouse"})
    assert d["sink"]["columns"][-2:] == ["UInt64", "UInt64"]
    # refusals: an update, a transformer chain
    s = sink.Sink(updateable=True)
    with pytest.raises(engine.EngineError) as ei:
        s.push(rows.RowsImage([ChangeItem(K.KIND_UPDATE, 0, [go.int32(1), go.string("a"), go.float64(1)], {0: go.int32(1)})], [("public", "t", schema)]))
    assert ei.value.rc == -2 and "Collapse" in str(ei.value)
    s.close()
    with pytest.raises(engine.EngineError):
        sink.Sink(transformers=[{"rename_tables": {"renameTables": []}}], updateable=True)


def test_serialize_to_string_reference_cases():
    """registry/to_string/to_string_test.go:130-168 (TestAllTypesToStringTransformer): the expected text of SerializeToString per type, checked on the
    host formatter table_splitter uses (through generated table names) and on the oracle's restatement — years before 0 and beyond 9999, nanoseconds,
    float32 / float64 shortest forms, nil as "<nil>"."""
    def secs(text):
        return int(np.datetime64(text, "s").astype(np.int64))
    cases = [("int64", go.int64(981274987), "981274987"), ("int32", go.int32(-12049182), "-12049182"), ("int16", go.int16(12313), "12313"), ("int8", go.int8(-14), "-14"),
             ("uint64", go.uint64(1142423562), "1142423562"), ("uint32", go.uint32(0), "0"), ("uint16", go.uint16(65212), "65212"), ("uint8", go.uint8(213), "213"),
             ("float", go.float32(123.123), "123.123"), ("double", go.float64(-12344.12334341), "-12344.12334341"),
             ("string", go.bytes(b"bytes"), "bytes"), ("utf8", go.string("string"), "string"), ("boolean", go.bool(True), "true"),
             ("date", go.time(secs("-1232-02-23T00:00:00")), "-1232-02-23"), ("date", go.time(secs("14124-01-12T00:00:00")), "14124-01-12"),
             ("datetime", go.time(secs("2311-12-01T01:02:04"), 5), "2311-12-01T01:02:04.000000005Z"),
             ("timestamp", go.time(secs("1231-05-23T09:08:07"), 6), "1231-05-23T09:08:07.000000006Z"),
             ("date", go.nil, "<nil>"), ("datetime", go.nil, "<nil>"), ("boolean", go.nil, "<nil>"), ("utf8", go.nil, "<nil>"), ("int64", go.nil, "<nil>")]
    for typ, val, want in cases:
        assert mo.serialize_to_string(val, typ) == want, (typ, val)
        s = sink.Sink(transformers=[{"table_splitter": {"columns": ["c"], "splitter": "|"}}])
        s.push(rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, [val])], [("", "", [{"name": "c", "type": typ}])]))
        assert [e["out"][1] for e in s.events] == [want], (typ, val)          # an empty original table name contributes no component
        s.close()
