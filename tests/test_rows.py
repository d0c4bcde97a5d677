"""Host transpose: []ChangeItem in row form -> columnar tf_batch and back (tfgpu_rows_to_batch / tfgpu_batch_to_rows).
The expectations are stated independently in numpy on the columnar side: a batch made by the workload generator is turned into ChangeItems
carrying the canonical Go types, flattened like the shim would, transposed by the product, and must come back as the same columns."""
import os
import numpy as np
import pytest

from transferia_b200 import abi, engine, rows, workload
from transferia_b200.rows import ChangeItem, go
from test_gpu_parity import all_types_batch


def _lens(col, n):
    o = np.asarray(col.offsets)
    return o.astype(np.int64) if col.lens_width else np.diff(o.astype(np.int64))


def _valid(col, n):
    return np.ones(n, bool) if col.validity is None else np.unpackbits(np.asarray(col.validity), bitorder="little")[:n].astype(bool)


def assert_same_batch(got: abi.Batch, want: abi.Batch):
    assert got.nrows == want.nrows and len(got.columns) == len(want.columns)
    n = want.nrows
    for k, (g, w) in enumerate(zip(got.columns, want.columns)):
        assert g.type == w.type, (k, g.type, w.type)
        if n == 0:
            continue
        vg, vw = _valid(g, n), _valid(w, n)
        assert (vg == vw).all(), k
        if w.type in abi.VAR_TYPES:
            lw = np.where(vw, _lens(w, n), 0)
            assert (_lens(g, n) == lw).all(), k
            off = np.asarray(w.offsets).astype(np.int64) if not w.lens_width else np.concatenate([[0], np.cumsum(_lens(w, n))])
            heap_w = np.asarray(w.heap).tobytes() if w.heap is not None else b""
            want_heap = b"".join(heap_w[off[r]:off[r + 1]] for r in range(n) if vw[r])
            assert (np.asarray(g.heap).tobytes() if g.heap is not None else b"") == want_heap, k
            if w.type == abi.TF_ANY and w.aux is not None:
                ga = np.zeros(n, np.uint8) if g.aux is None else np.asarray(g.aux)
                assert (ga[vw] == np.asarray(w.aux)[vw]).all(), k
        else:
            a, b = np.asarray(g.values)[vw], np.asarray(w.values)[vw]
            assert a.tobytes() == b.tobytes(), k               # bit-exact (NaN payloads, -0.0)
            if w.type in abi.TIME_TYPES:
                na = np.zeros(n, np.uint32) if g.aux is None else np.asarray(g.aux)
                nb = np.zeros(n, np.uint32) if w.aux is None else np.asarray(w.aux)
                assert (na[vw] == nb[vw]).all(), k


def test_hits_rows_transpose_equals_generator_columns():
    batch, schema = workload.make_hits_batch(3000, seed=4)
    img = rows.RowsImage(rows.items_from_batch(batch), [("public", "hits", schema)])
    pool = rows.Columnar()
    for threads in (1, 4):
        t = pool.rows_to_batch(img, threads=threads)
        assert_same_batch(t.batch, batch)
        # short cells travel as uint8 lengths (TF_COL_LENS8), exactly what Batch.narrow() builds on the Python side
        nb = batch.narrow()
        assert [c.lens_width for c in t.batch.columns] == [c.lens_width for c in nb.columns]
        assert t.batch.kinds is None and t.old is None
    pool.close()


def test_all_types_with_nulls_round_trip_through_the_inverse():
    batch, schema = all_types_batch(3000, seed=11)
    image, off = rows.batch_to_rows(batch)                                # columnar -> ColumnValues images (the inverse)
    items = rows.items_from_batch(batch)
    want = bytearray()
    for it in items[:50]:
        for v in it.values: rows.encode_value(want, v)
    assert image[:int(off[50])] == bytes(want)                            # same image the Python stand-in of the shim writes
    # feed the product's own images back: items whose values are already encoded
    class Raw(rows.RowsImage):
        pass
    img = rows.RowsImage(items, [("db", "t", schema)])
    assert img._vals[:img.values_len].tobytes() == image
    pool = rows.Columnar()
    t = pool.rows_to_batch(img, threads=3)
    assert_same_batch(t.batch, batch)
    assert 2 in [c.lens_width for c in t.batch.columns] and 1 in [c.lens_width for c in t.batch.columns]   # 20000-byte cells -> uint16 lengths
    pool.close()


def test_loose_value_types_become_loose_columns():
    schema = [{"name": "a", "type": "int32"}, {"name": "b", "type": "uint8"}, {"name": "c", "type": "double"}, {"name": "d", "type": "double"},
              {"name": "e", "type": "int64"}, {"name": "f", "type": "any"}, {"name": "g", "type": "datetime"}, {"name": "h", "type": "float"}]
    items = [ChangeItem(values=[go.int64(7), go.int64(200), go.float32(1.5), go.number("0.1"), go.int8(-3), go.int64(12), go.int64(1700000000), go.float32(2.5)]),
             ChangeItem(values=[go.int64(-2 ** 40), go.int64(-1), go.float32(-0.0), go.number("1e400"), go.int64(2 ** 62), go.bool(True), go.int64(5), go.nil]),
             ChangeItem(values=[go.nil, go.int64(3), go.float64(2.25), go.nil, go.bool(False), go.json('{"a":[1,2]}'), go.nil, go.float32(-1)])]
    pool = rows.Columnar()
    t = pool.rows_to_batch(rows.RowsImage(items, [("", "t", schema)]))
    c = t.batch.columns
    assert [x.type for x in c] == [abi.TF_INT64, abi.TF_INT64, abi.TF_DOUBLE, abi.TF_DOUBLE, abi.TF_INT64, abi.TF_ANY, abi.TF_INT64, abi.TF_FLOAT]
    assert list(np.asarray(c[0].values)[:2]) == [7, -2 ** 40] and list(_valid(c[0], 3)) == [True, True, False]
    assert list(np.asarray(c[1].values)) == [200, -1, 3]                  # the device's Strictify raises the range / cast errors (strictify.go:159-181)
    assert np.asarray(c[2].values).tobytes() == np.array([1.5, -0.0, 2.25]).tobytes()          # float32 widened exactly
    assert list(np.asarray(c[3].values)[:2]) == [0.1, np.inf]            # json.Number text -> ParseFloat
    assert list(np.asarray(c[4].values)) == [-3, 2 ** 62, 0]              # int8 / int64 / bool mixed: one signed class, widened
    assert np.asarray(c[5].heap).tobytes() == b'12true{"a":[1,2]}' and list(np.asarray(c[5].offsets)) == [2, 4, 11]
    assert list(np.asarray(c[6].values)[:2]) == [1700000000, 5]
    # refusals: the batch stays on the Go path
    for bad in ([go.string("12")] + [go.nil] * 7, [go.int64(1), go.uint64(2)] + [go.nil] * 6):
        its = [ChangeItem(values=bad), ChangeItem(values=[go.int64(1), go.int64(2)] + [go.nil] * 6)]
        with pytest.raises(engine.EngineError) as ei:
            pool.rows_to_batch(rows.RowsImage(its, [("", "t", schema)]))
        assert ei.value.rc == -2
    with pytest.raises(engine.EngineError) as ei:                         # a row with too few values and no TF_ITEM_SPARSE
        pool.rows_to_batch(rows.RowsImage([ChangeItem(values=[go.int64(1)])], [("", "t", schema)]))
    assert ei.value.rc == -3
    pool.close()


def test_kinds_meta_sparse_rows_and_old_keys():
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "name", "type": "utf8"}, {"name": "ts", "type": "timestamp"}]
    other = [{"name": "x", "type": "int8"}]
    items = [
        ChangeItem(rows.KIND_INIT_TABLE_LOAD, 0),
        ChangeItem(rows.KIND_INSERT, 0, [go.int32(1), go.string("a"), go.time(10, 5)], id=7, lsn=100, commit_time=123456789, txid=b"tx-1"),
        ChangeItem(rows.KIND_INSERT, 1, [go.int8(9)]),
        ChangeItem(rows.KIND_UPDATE, 0, {0: go.int32(2), 1: go.string("bb")}, {0: go.int32(1)}, id=8, lsn=101, commit_time=2, txid=b"tx-22"),
        ChangeItem(rows.KIND_DELETE, 0, {}, {0: go.int32(2), 1: go.string("bb")}, id=9, lsn=102, commit_time=3),
        ChangeItem(rows.KIND_DONE_TABLE_LOAD, 0),
    ]
    img = rows.RowsImage(items, [("public", "users", schema), ("public", "other", other)])
    pool = rows.Columnar()
    t = pool.rows_to_batch(img, table=0)                                  # the row events of table 0, in order: items 1, 3, 4
    b = t.batch
    assert b.nrows == 3 and list(np.asarray(b.kinds)) == [0, 1, 2]
    assert list(t.ids) == [7, 8, 9] and list(t.lsn) == [100, 101, 102] and list(t.commit_time) == [123456789, 2, 3]
    m = t.meta.contents
    assert rows._view(m.txid_heap, 9).tobytes() == b"tx-1tx-22" and list(rows._view(m.txid_offsets, 16, np.uint32)) == [0, 4, 9, 9]
    assert list(np.asarray(b.columns[0].values)) == [1, 2, 0] and list(_valid(b.columns[0], 3)) == [True, True, False]
    assert np.asarray(b.columns[1].heap).tobytes() == b"abb" and list(_valid(b.columns[2], 3)) == [True, False, False]
    assert list(np.asarray(b.columns[2].aux)) == [5, 0, 0]
    assert list(t.old_present) == [1, 1, 0] and list(t.old_row_has) == [0, 1, 1]
    ob = t.old_batch
    assert list(np.asarray(ob.columns[0].values)) == [0, 1, 2] and list(_valid(ob.columns[0], 3)) == [False, True, True]
    assert np.asarray(ob.columns[1].heap).tobytes() == b"bb" and list(_valid(ob.columns[1], 3)) == [False, False, True]
    assert ob.columns[2].values is None and ob.columns[2].offsets is None
    t1 = pool.rows_to_batch(img, table=1)
    assert t1.batch.nrows == 1 and list(np.asarray(t1.batch.columns[0].values)) == [9] and t1.old is None
    # an explicit item list must name row events of that table
    with pytest.raises(engine.EngineError):
        pool.rows_to_batch(img, table=0, item_idx=[0])
    assert pool.rows_to_batch(img, table=0, item_idx=[4, 1]).batch.nrows == 2
    pool.close()


@pytest.mark.gpu
def test_transposed_rows_through_the_device_equal_the_oracle(eng, po):
    """Rows -> transpose (loose int64 values in the narrow integer columns, as a JSON-decoding source hands them) -> device Strictify +
    filter + cast + native block == the oracle on the strict columnar batch."""
    batch, schema = workload.make_hits_batch(4000, seed=6)
    items = rows.items_from_batch(batch)
    loose = {k for k, c in enumerate(batch.columns) if c.type in (abi.TF_INT16, abi.TF_INT32)}
    for it in items:
        it.values = [go.int64(v[1]) if k in loose and v[0] != rows.V_NIL else v for k, v in enumerate(it.values)]
    pool = rows.Columnar()
    t = pool.rows_to_batch(rows.RowsImage(items, [("public", "hits", schema)]))
    assert all(t.batch.columns[k].type == abi.TF_INT64 for k in loose)
    trs = workload.headline_transformers(workload.counterid_threshold(batch, schema))
    pid = eng.plan("public", "hits", schema, trs, {"type": "clickhouse"})
    got = eng.push_encode(pid, t.batch, abi.TF_WIRE_CH_NATIVE)
    want = po.push_encode(batch, po.build_plan("public", "hits", schema, trs), abi.TF_WIRE_CH_NATIVE)
    assert got.rows_out == want.rows_out and got.wire == want.raw and not got.errors
    pool.close()


def _numpy_select(batch: abi.Batch, keep: np.ndarray) -> abi.Batch:
    """The expectation for tfgpu_batch_gather, stated with numpy fancy indexing over plain uint32 offsets."""
    idx = np.nonzero(keep)[0]; n = batch.nrows; cols = []
    for c in batch.columns:
        val = None if c.validity is None else abi.pack_validity(np.unpackbits(np.asarray(c.validity), bitorder="little")[:n].astype(bool)[idx])
        aux = None if c.aux is None else np.asarray(c.aux)[idx]
        if c.type in abi.VAR_TYPES:
            ln = _lens(c, n); off = np.concatenate([[0], np.cumsum(ln)]); heap = np.asarray(c.heap).tobytes()
            cells = [heap[off[r]:off[r + 1]] for r in idx]
            o = np.zeros(len(idx) + 1, np.uint32); np.cumsum([len(x) for x in cells], out=o[1:])
            cols.append(abi.Column(c.type, None, val, o, np.frombuffer(b"".join(cells), np.uint8), aux))
        else:
            cols.append(abi.Column(c.type, np.asarray(c.values)[idx], val, None, None, aux))
    return abi.Batch(len(idx), cols, None if batch.kinds is None else np.asarray(batch.kinds)[idx])


def test_host_gather_equals_numpy_selection():
    pool = rows.Columnar()
    rng = np.random.default_rng(3)
    for make, n in ((lambda: all_types_batch(70_000, seed=2)[0], 70_000), (lambda: workload.make_hits_batch(40_000, seed=8)[0].narrow(), 40_000)):
        batch = make()
        batch.kinds = rng.integers(0, 3, n).astype(np.uint8)
        for p in (0.28, 0.0, 1.0):
            keep = (rng.random(n) < p).astype(np.uint8)
            for threads in (1, 5):
                got, sel = pool.gather(batch, keep, threads)
                want = _numpy_select(batch, keep)
                assert (sel == np.nonzero(keep)[0]).all()
                assert_same_batch(got, want)
                assert (np.asarray(got.kinds) == np.asarray(want.kinds)).all() if want.nrows else True
                assert [c.lens_width for c in got.columns] == [c.lens_width for c in batch.columns]     # the layout of the input is kept
    pool.close()


def _py_transpose(items, schema):
    """The documented transposer rules (include/tfgpu_sink.h, tfgpu_rows_to_batch) restated over Python lists: per column the physical type
    (schema type when every value carries the canonical tag; the single foreign numeric type; else INT64 / UINT64 / DOUBLE by class), values,
    validity, text cells. Returns [(phys_tf, values list with None for nil)] or raises ValueError where the product refuses."""
    R = rows
    canon = {abi.TF_INT8: R.V_INT8, abi.TF_INT16: R.V_INT16, abi.TF_INT32: R.V_INT32, abi.TF_INT64: R.V_INT64, abi.TF_UINT8: R.V_UINT8, abi.TF_UINT16: R.V_UINT16,
             abi.TF_UINT32: R.V_UINT32, abi.TF_UINT64: R.V_UINT64, abi.TF_FLOAT: R.V_FLOAT32, abi.TF_DOUBLE: R.V_FLOAT64, abi.TF_BOOLEAN: R.V_BOOL,
             abi.TF_INTERVAL: R.V_DURATION, abi.TF_DATE: R.V_TIME, abi.TF_DATETIME: R.V_TIME, abi.TF_TIMESTAMP: R.V_TIME}
    tag_tf = {R.V_BOOL: abi.TF_BOOLEAN, R.V_INT8: abi.TF_INT8, R.V_INT16: abi.TF_INT16, R.V_INT32: abi.TF_INT32, R.V_INT64: abi.TF_INT64, R.V_UINT8: abi.TF_UINT8,
              R.V_UINT16: abi.TF_UINT16, R.V_UINT32: abi.TF_UINT32, R.V_UINT64: abi.TF_UINT64, R.V_FLOAT32: abi.TF_FLOAT, R.V_FLOAT64: abi.TF_DOUBLE, R.V_JSONNUM: abi.TF_DOUBLE}
    signed, unsigned, floats = {R.V_BOOL, R.V_INT8, R.V_INT16, R.V_INT32, R.V_INT64}, {R.V_UINT8, R.V_UINT16, R.V_UINT32, R.V_UINT64}, {R.V_FLOAT32, R.V_FLOAT64, R.V_JSONNUM}
    out = []
    for c, col in enumerate(schema):
        tf = abi.YT_NAME_TO_TF[col["type"]]
        cells = []
        for it in items:
            v = it.values.get(c, go.nil) if isinstance(it.values, dict) else it.values[c]
            cells.append(v)
        tags = {v[0] for v in cells if v[0] != R.V_NIL}
        if tf in abi.VAR_TYPES:
            if tf != abi.TF_ANY and tags - {R.V_STRING, R.V_BYTES}: raise ValueError("non-text in a string column")
            if tf == abi.TF_ANY and tags & (floats - {R.V_JSONNUM} | {R.V_TIME, R.V_DURATION, R.V_BYTES}): raise ValueError("float / time inside any")
            def text(v):
                if v[0] == R.V_NIL: return None
                if v[0] in (R.V_STRING, R.V_BYTES, R.V_JSON, R.V_JSONNUM): return bytes(v[1])
                if v[0] == R.V_BOOL: return b"true" if v[1] else b"false"
                return str(int(v[1])).encode()
            out.append((tf, [text(v) for v in cells])); continue
        if not tags or tags == {canon[tf]} or (tf == abi.TF_DOUBLE and tags == {R.V_JSONNUM}): phys = tf
        elif R.V_TIME in tags and tf in abi.TIME_TYPES: raise ValueError("time mixed")
        elif tags - (signed | unsigned | floats | {R.V_DURATION}): raise ValueError("text in a fixed column")
        elif len(tags) == 1: t = next(iter(tags)); phys = abi.TF_INT64 if t == R.V_DURATION else tag_tf[t]
        elif not (tags - signed) or not (tags - (signed | {R.V_DURATION})): phys = abi.TF_INT64
        elif not (tags - unsigned): phys = abi.TF_UINT64
        elif not (tags - floats): phys = abi.TF_DOUBLE
        else: raise ValueError("mixed classes")
        def num(v):
            if v[0] == R.V_NIL: return None
            if v[0] == R.V_TIME: return v[1]
            if v[0] == R.V_JSONNUM: return float(v[1])
            if v[0] == R.V_FLOAT32: return float(np.float32(v[1]))
            return v[1]
        out.append((phys, [num(v) for v in cells]))
    return out


def test_transposer_fuzz_against_the_documented_rules():
    rng = np.random.default_rng(77)
    types = ["int8", "int32", "int64", "uint16", "uint64", "float", "double", "boolean", "utf8", "string", "any", "date", "timestamp", "interval"]
    pool = rows.Columnar()
    agreed = refused = 0
    for rnd in range(120):
        ncol = int(rng.integers(1, 9)); n = int(rng.integers(1, 400))
        schema = [{"name": f"c{k}", "type": str(rng.choice(types))} for k in range(ncol)]
        def value(tf, mode):
            if rng.random() < 0.12: return go.nil
            if tf in (abi.TF_UTF8, abi.TF_BYTES):
                b = bytes(rng.integers(0, 256, int(rng.integers(0, 300 if mode else 20)), dtype=np.uint8))
                return go.string(b) if rng.random() < 0.7 else go.bytes(b)
            if tf == abi.TF_ANY:
                k = rng.integers(0, 4); return [go.json(b'{"a":1}'), go.string("s"), go.int64(int(rng.integers(-5, 5))), go.bool(rng.random() < 0.5)][int(k)]
            if tf in abi.TIME_TYPES: return go.time(int(rng.integers(-10**9, 10**10)), int(rng.integers(0, 10**9)) if mode else 0)
            if tf == abi.TF_INTERVAL: return go.duration(int(rng.integers(-10**12, 10**12)))
            if tf == abi.TF_BOOLEAN and not mode: return go.bool(rng.random() < 0.5)
            if tf in (abi.TF_FLOAT, abi.TF_DOUBLE):
                if mode == 0: return go.float32(float(rng.standard_normal())) if tf == abi.TF_FLOAT else go.float64(float(rng.standard_normal()))
                return [go.float32(1.5), go.float64(2.25), go.number("0.125")][int(rng.integers(0, 3))]
            canon = {abi.TF_INT8: go.int8, abi.TF_INT32: go.int32, abi.TF_INT64: go.int64, abi.TF_UINT16: go.uint16, abi.TF_UINT64: go.uint64, abi.TF_BOOLEAN: go.bool}[tf]
            if mode == 0: return canon(int(rng.integers(0, 100)))
            if mode == 1: return go.int64(int(rng.integers(-1000, 1000)))                      # one foreign type
            return [go.int8, go.int16, go.int64, go.bool][int(rng.integers(0, 4))](int(rng.integers(0, 2)))   # mixed signed widths
        modes = [int(rng.integers(0, 3)) for _ in range(ncol)]
        if rng.random() < 0.1: modes[0] = 9                                                       # provoke a refusal
        items = []
        for r in range(n):
            vals = [value(abi.YT_NAME_TO_TF[c["type"]], m) if m != 9 else (go.string("x") if abi.YT_NAME_TO_TF[c["type"]] not in abi.VAR_TYPES else go.int64(1)) for c, m in zip(schema, modes)]
            if rng.random() < 0.15:
                keep = [k for k in range(ncol) if rng.random() < 0.7]
                items.append(ChangeItem(rows.KIND_UPDATE, 0, {k: vals[k] for k in keep}))
            else:
                items.append(ChangeItem(rows.KIND_INSERT, 0, vals))
        try:
            want = _py_transpose(items, schema)
        except ValueError:
            want = None
        try:
            got = pool.rows_to_batch(rows.RowsImage(items, [("d", "t", schema)]), threads=int(rng.integers(1, 5)))
        except engine.EngineError as ex:
            assert want is None and ex.rc == -2, (rnd, schema, modes)
            refused += 1; continue
        assert want is not None, (rnd, schema, modes)
        agreed += 1
        b = got.batch
        for c, (phys, cells) in enumerate(want):
            col = b.columns[c]
            assert col.type == phys, (rnd, c, col.type, phys)
            valid = _valid(col, n)
            assert list(valid) == [v is not None for v in cells], (rnd, c)
            if phys in abi.VAR_TYPES:
                ln = _lens(col, n); heap = np.asarray(col.heap).tobytes() if col.heap is not None else b""
                assert heap == b"".join(v for v in cells if v is not None) and [int(x) for x in ln] == [0 if v is None else len(v) for v in cells], (rnd, c)
            elif phys in abi.TIME_TYPES:
                vals = np.asarray(col.values); ns = np.zeros(n, np.uint32) if col.aux is None else np.asarray(col.aux)
                assert [(int(vals[r]), int(ns[r])) for r in range(n) if valid[r]] == [v for v in cells if v is not None], (rnd, c)
            else:
                vals = np.asarray(col.values)
                wantv = np.array([0 if v is None else v for v in cells]).astype(abi.FIXED_DTYPE[phys])
                assert vals[valid].tobytes() == wantv[valid].tobytes(), (rnd, c, phys)
    assert agreed > 60 and refused > 3, (agreed, refused)
    pool.close()


def test_two_pools_in_two_threads_do_not_interfere():
    import threading
    batch, schema = workload.make_hits_batch(20_000, seed=31)
    img = rows.RowsImage(rows.items_from_batch(batch.slice(0, 3000)), [("public", "hits", schema)])
    keep = (np.arange(20_000) % 3 == 0).astype(np.uint8)
    errs = []
    def work(k):
        try:
            pool = rows.Columnar()
            for _ in range(6):
                assert_same_batch(pool.rows_to_batch(img, threads=3).batch, batch.slice(0, 3000))
                g, sel = pool.gather(batch, keep, 3)
                assert g.nrows == int(keep.sum()) and (sel == np.nonzero(keep)[0]).all()
            pool.close()
        except BaseException as ex:   # noqa: BLE001
            errs.append(ex)
    ths = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert not errs, errs


def test_transposer_rejects_malformed_images_without_crashing():
    """Truncated / random value images, offsets past the end, column indexes outside the schema, wild value counts: an error code, never a
    read past the image (the image ends right at the buffer's end here, so an overrun would fault under the allocator's guard more often than not)."""
    import ctypes as C
    rng = np.random.default_rng(5)
    schema = [{"name": "a", "type": "int32"}, {"name": "s", "type": "utf8"}, {"name": "t", "type": "timestamp"}]
    good = rows.RowsImage([ChangeItem(rows.KIND_INSERT, 0, [go.int32(1), go.string("abc"), go.time(5, 6)], old_keys={0: go.int32(0)}) for _ in range(50)], [("d", "t", schema)])
    base = good._vals[:good.values_len].copy()
    pool = rows.Columnar(); outcomes = set()
    for trial in range(300):
        vals = base.copy()
        kind = trial % 6
        if kind == 0: vals = vals[: int(rng.integers(0, len(vals)))]                                  # truncated
        elif kind == 1: vals[rng.integers(0, len(vals), 8)] = rng.integers(0, 256, 8)                  # flipped bytes (tags, lengths, indexes)
        elif kind == 2: vals = rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8)         # noise
        img = rows.RowsImage([], [("d", "t", schema)])
        items = (rows.TfItem * 50)()
        for r in range(50):
            C.memmove(C.byref(items[r]), C.byref(good._items[r]), C.sizeof(rows.TfItem))
            if kind == 3: items[r].values_off = int(rng.integers(0, 2 ** 40))
            if kind == 4: items[r].n_values = int(rng.integers(0, 2 ** 31))
            if kind == 5: items[r].old_keys_off = int(rng.integers(0, 2 ** 20))
        buf = np.ascontiguousarray(vals)
        img.struct.n_items = 50; img.struct.items = C.cast(items, C.POINTER(rows.TfItem)); img.struct.values = buf.ctypes.data if len(buf) else None; img.struct.values_len = len(buf)
        try:
            pool.rows_to_batch(img, threads=2); outcomes.add(0)
        except engine.EngineError as ex:
            assert ex.rc in (-2, -3), ex.rc; outcomes.add(ex.rc)
    assert -3 in outcomes
    pool.close()


def test_strict_single_decode_path_equals_the_general_path():
    """Strictly typed value lists take the one-decode path of the transposer (fixed-width values written while decoding, var-width cells
    copied from noted image offsets); TFGPU_TRANSPOSE_GENERAL forces the two-pass path that also lays out loose columns. Same columns
    byte for byte: dense and sparse rows, nils, nanoseconds, raw strings inside `any`, empty and > 64 KiB cells, several chunk sizes."""
    rng = np.random.default_rng(99)
    batch, schema = all_types_batch(5000, seed=23)
    items = rows.items_from_batch(batch)
    for k in range(0, len(items), 7):                                     # every 7th row carries a column subset (absent = nil)
        keep = sorted(rng.choice(len(schema), size=int(rng.integers(0, len(schema))), replace=False).tolist())
        items[k] = ChangeItem(rows.KIND_UPDATE, 0, {c: items[k].values[c] for c in keep}, {0: items[k].values[0]})
    any_col = [i for i, c in enumerate(schema) if c["type"] == "any"]
    utf_col = [i for i, c in enumerate(schema) if c["type"] == "utf8"]
    for k in range(3, len(items), 11):
        if isinstance(items[k].values, list):
            for c in any_col: items[k].values[c] = go.string("raw text %d" % k)
            for c in utf_col[:1]: items[k].values[c] = go.string("x" * (70000 if k % 5 == 0 else 0))
    img = rows.RowsImage(items, [("db", "t", schema)])
    pool = rows.Columnar()

    def snapshot(t):
        out = []
        for c in t.batch.columns:
            out.append((c.type, c.lens_width, *(None if a is None else np.asarray(a).tobytes() for a in (c.values, c.validity, c.offsets, c.heap, c.aux))))
        return out, None if t.batch.kinds is None else np.asarray(t.batch.kinds).tobytes()
    try:
        for threads in (1, 3, 8):
            os.environ.pop("TFGPU_TRANSPOSE_GENERAL", None)
            strict = snapshot(pool.rows_to_batch(img, threads=threads))
            os.environ["TFGPU_TRANSPOSE_GENERAL"] = "1"
            general = snapshot(pool.rows_to_batch(img, threads=threads))
            assert len(strict[0]) == len(general[0]) and strict[1] == general[1]
            for c, (a, b) in enumerate(zip(strict[0], general[0])):
                assert a[:2] == b[:2], (c, a[:2], b[:2])
                for name, x, y in zip(("values", "validity", "offsets", "heap", "aux"), a[2:], b[2:]):
                    if name == "validity" and x is not None and y is not None:
                        n = len(items); assert x[:n // 8] == y[:n // 8] and (n % 8 == 0 or x[n // 8] == y[n // 8]), (c, name)
                    elif name == "heap" and x is not None and y is not None:
                        assert x == y, (c, name)
                    else:
                        assert (x is None) == (y is None) and (x is None or x == y), (c, name, threads)
    finally:
        os.environ.pop("TFGPU_TRANSPOSE_GENERAL", None)
    pool.close()


def test_inverse_transposer_in_parallel_equals_the_sequential_walk(monkeypatch):
    """tfgpu_batch_to_rows measures and writes its row ranges on several threads from 8192 rows on (TFGPU_INVERSE_SEQUENTIAL forces one walk
    over all rows): same image and offsets, for uint32 offsets and for narrow length arrays, with nulls in every column."""
    batch, _schema = all_types_batch(20000, seed=31)
    for b in (batch, batch.narrow()):
        monkeypatch.delenv("TFGPU_INVERSE_SEQUENTIAL", raising=False)
        image, off = rows.batch_to_rows(b)
        monkeypatch.setenv("TFGPU_INVERSE_SEQUENTIAL", "1")
        image_s, off_s = rows.batch_to_rows(b)
        assert image == image_s and (off == off_s).all() and int(off[-1]) == len(image)
    monkeypatch.delenv("TFGPU_INVERSE_SEQUENTIAL", raising=False)
    pool = rows.Columnar()                                                       # and back: the transposer reads what the inverse wrote
    import ctypes as C
    n = batch.nrows
    image, off = rows.batch_to_rows(batch)
    items = (rows.TfItem * n)()
    for r in range(n):
        items[r].values_off = int(off[r]); items[r].n_values = len(_schema); items[r].old_keys_off = rows.NO_OLD_KEYS
    img = rows.RowsImage([], [("db", "t", _schema)])
    vals = np.frombuffer(image, dtype=np.uint8).copy()
    img.struct.n_items = n; img.struct.items = C.cast(items, C.POINTER(rows.TfItem)); img.struct.values = vals.ctypes.data; img.struct.values_len = len(image)
    assert_same_batch(pool.rows_to_batch(img, threads=4).batch, batch)
    pool.close()
