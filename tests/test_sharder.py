"""sharder_transformer (pkg/transformer/registry/sharder/sharder.go): ChangeItem.PartID =
decimal(CRC32-IEEE(join(".", SerializeToString(value) of the matched columns)) % ShardsNum).

The oracle is pinned by the six part ids of the reference's canon file (sharder/gotest/canondata/result.json, produced by
sharder_test.go:TestSharderTransformer); the device kernel is compared with the oracle."""
import zlib

import numpy as np
import pytest

from transferia_b200 import abi

# sharder_test.go:86-112: the three items, as (Go value kind, column type) per column
CANON = {
    # transformer -> [(table, expected part id)]  (canondata/result.json "part")
    "all_cols_2": [("table1", 1), ("table2", 0), ("a_table3", 0)],
    "exclude_column2_4": [("table1", 2), ("table2", 3)],
    "include_column1_column3_8": [("a_table3", 4)],
}


def _items(po):
    V, S = po.make_val, lambda v, t: po.serialize_to_string(v, abi.YT_NAME_TO_TF[t])
    sec_1703 = -8425641600          # 1703-01-02T00:00:00Z
    return {
        "table1": {"column1": S(V(po.OG_STRING, s=b"value1"), "string"), "column2": S(V(po.OG_INT64, i=123), "int64"),
                   "column3": S(V(po.OG_INT32, i=1234), "int32"), "column4": S(V(po.OG_BOOL, i=1), "boolean")},
        # the item names its first column "colunm1" (sic): the schema's column1 has no value -> nil
        "table2": {"column1": S(V(po.OG_NIL), "string"), "column2": S(V(po.OG_TIME, i=sec_1703), "date"),
                   "column3": S(V(po.OG_FLOAT64, f=123.123), "double"), "column4": S(V(po.OG_FLOAT32, f=float(np.float32(312.321))), "float")},
        "a_table3": {"column2": S(V(po.OG_INT8, i=-3), "int8"), "column3": S(V(po.OG_UINT32, u=12345), "uint32"),
                     "column4": S(V(po.OG_DURATION, i=60 * 10**9), "date")},
    }


def test_oracle_reproduces_reference_part_ids(po):
    it = _items(po)
    assert it["table2"]["column2"] == b"1703-01-02" and it["a_table3"]["column4"] == b"1m0s" and it["table2"]["column1"] == b"<nil>"
    pick = {"all_cols_2": (lambda n: True, 2), "exclude_column2_4": (lambda n: n != "column2", 4), "include_column1_column3_8": (lambda n: n in ("column1", "column3"), 8)}
    for tr, cases in CANON.items():
        match, shards = pick[tr]
        for table, want in cases:
            joined = b".".join(v for k, v in it[table].items() if match(k))
            assert po.crc32_ieee(joined) == zlib.crc32(joined)
            assert po.crc32_ieee(joined) % shards == want, (tr, table)


def test_oracle_sharder_in_the_chain(po):
    """The step reads the values as the transformers before it left them; Suitable follows sharder.go:93-105."""
    schema = [{"name": "column1", "type": "utf8", "key": True}, {"name": "column2", "type": "int64"}, {"name": "column3", "type": "int32"}, {"name": "column4", "type": "boolean"}]
    b = abi.Batch(2, [abi.strings_to_column(abi.TF_UTF8, [b"value1", None]), abi.fixed_to_column(abi.TF_INT64, [123, 5]),
                      abi.fixed_to_column(abi.TF_INT32, [1234, 6]), abi.fixed_to_column(abi.TF_BOOLEAN, [1, 0])])
    sh = lambda **kw: {"sharder_transformer": dict({"shardsCount": "2"}, **kw)}
    plan = po.build_plan("db", "table1", schema, [sh()])
    assert list(po.shard_ids(b, plan)) == [1, zlib.crc32(b"<nil>.5.6.false") % 2]                       # item1 of the canon data -> part 1
    plan = po.build_plan("db", "table1", schema, [sh(shardsCount="4", columns={"excludeColumns": ["column2"]})])
    assert po.shard_ids(b, plan)[0] == 2
    assert po.build_plan("db", "table1", schema, [sh(columns={"includeColumns": ["nope"]})]).steps == []   # not Suitable
    assert po.build_plan("db", "table1", schema, [sh(tables={"includeTables": ["^db.other$"]})]).steps == []
    mask = {"mask_field": {"columns": ["column1"], "maskFunctionHash": {"userDefinedSalt": "s"}}}
    digest = po.hmac_hex(b"s", b"value1").encode()
    assert po.shard_ids(b, po.build_plan("db", "table1", schema, [mask, sh(shardsCount="1000")]))[0] == zlib.crc32(digest + b".123.1234.true") % 1000
    assert po.shard_ids(b, po.build_plan("db", "table1", schema, [sh(shardsCount="1000"), mask]))[0] == zlib.crc32(b"value1.123.1234.true") % 1000
    # the last sharder wins; rows a filter drops get no id
    two = [sh(shardsCount="7"), {"filter_rows": {"filter": "column2 > 100"}}, sh(shardsCount="1000", columns={"includeColumns": ["column3"]})]
    assert list(po.shard_ids(b, po.build_plan("db", "table1", schema, two))) == [zlib.crc32(b"1234") % 1000]


def test_product_plan_sharder_suitable_and_refusals(po):
    """The product's host-side plan builder (libtfgpu.so, no GPU needed) against sharder_test.go:61-83 (Suitable per table / column
    filter) and against the oracle's plan; configurations the device does not take are refused at plan time."""
    from transferia_b200 import engine
    t1 = [{"name": "column1", "type": "string", "key": True}, {"name": "column2", "type": "int64"}, {"name": "column3", "type": "int32"}, {"name": "column4", "type": "boolean"}]
    t3 = [{"name": "column2", "type": "int8"}, {"name": "column3", "type": "uint32"}, {"name": "column4", "type": "date"}]
    all_cols = {"sharder_transformer": {"shardsCount": "2"}}
    excl = {"sharder_transformer": {"shardsCount": "4", "tables": {"includeTables": ["db.table"]}, "columns": {"excludeColumns": ["column2"]}}}
    incl = {"sharder_transformer": {"shardsCount": "8", "tables": {"includeTables": ["db.a_table3"]}, "columns": {"includeColumns": ["column1", "column3"]}}}
    cases = [(all_cols, "table1", t1, True), (all_cols, "a_table3", t3, True), (excl, "table1", t1, True), (excl, "a_table3", t3, False),
             (incl, "table1", t1, False), (incl, "a_table3", t3, True)]
    for tr, table, schema, suitable in cases:
        d = engine.plan_validate("db", table, schema, [tr])
        plan = po.build_plan("db", table, schema, [tr])
        assert (len(d["steps"]) == 1) == suitable == (len(plan.steps) == 1), (tr, table)
        if suitable:
            assert d["steps"][0]["type"] == "sharder_transformer" and d["steps"][0]["cols"] == plan.steps[0]["cols"] and d["steps"][0]["shards"] == plan.steps[0]["shards"]
    assert engine.plan_validate("db", "a_table3", t3, [incl])["steps"][0]["cols"] == [1]          # only column3 exists there
    for bad in ({"shardsCount": "2", "is_random": True}, {"shardsCount": "x"}, {"shardsCount": ""}, {"shardsCount": "0"}, {"shardsCount": "4294967296"}):
        with pytest.raises(engine.EngineError):
            engine.plan_validate("db", "table1", t1, [{"sharder_transformer": bad}])
    with pytest.raises(engine.EngineError):      # number_to_float rewrites the `any` text the sharder would read
        engine.plan_validate("db", "t", [{"name": "a", "type": "any"}], [{"sharder_transformer": {"shardsCount": "2"}}, {"number_to_float_transformer": {}}])


@pytest.mark.gpu
def test_device_sharder_equals_oracle(eng, po):
    from test_gpu_parity import all_types_batch
    batch, schema = all_types_batch(3000, seed=21)
    sh = lambda **kw: {"sharder_transformer": dict({"shardsCount": "64"}, **kw)}
    mask = {"mask_field": {"columns": ["c_utf8", "n_int64"], "maskFunctionHash": {"userDefinedSalt": "salt"}}}
    chains = [[sh()], [sh(shardsCount="4294967295", columns={"includeColumns": ["^c_", "n_double", "n_any"]})],
              [mask, sh(columns={"includeColumns": ["c_utf8", "n_int64", "c_int8"]})], [sh(columns={"includeColumns": ["c_utf8", "n_int64"]}), mask],
              [{"filter_rows": {"filter": "c_int32 > 0"}}, {"convert_to_datetime": {"columns": {"includeColumns": ["c_uint32", "n_int32"]}}}, sh(shardsCount="1000")],
              [{"convert_to_string": {"columns": {"includeColumns": ["n_double", "n_timestamp", "c_utf8"]}}}, sh(shardsCount="3", columns={"excludeColumns": ["n_any"]})],
              [sh(shardsCount="9"), {"filter_columns": {"columns": {"excludeColumns": ["c_date"]}}}, sh(shardsCount="5")]]
    for trs in chains:
        pid = eng.plan("db", "t", schema, trs); plan = po.build_plan("db", "t", schema, trs)
        want = po.shard_ids(batch, plan)
        got_b, errs = eng.push_columns(pid, batch)
        assert got_b.nrows == len(want), trs
        assert eng.last_part_ids is not None and list(eng.last_part_ids) == list(want), trs
        r = eng.push_encode(pid, batch, 4)                  # serializer JSON rows carry the same ids
        assert list(r.part_ids) == list(want), trs
    pid = eng.plan("db", "t", schema, [])
    assert eng.push_encode(pid, batch, 4).part_ids is None


@pytest.mark.gpu
def test_device_sharder_refusals(eng):
    from transferia_b200.engine import EngineError
    schema = [{"name": "a", "type": "any"}, {"name": "i", "type": "int32"}]
    for trs in ([{"sharder_transformer": {"shardsCount": "2", "is_random": True}}], [{"sharder_transformer": {"shardsCount": "x"}}], [{"sharder_transformer": {"shardsCount": "0"}}],
                [{"sharder_transformer": {"shardsCount": "2"}}, {"number_to_float_transformer": {}}]):
        with pytest.raises(EngineError):
            eng.plan("db", "t", schema, trs)
