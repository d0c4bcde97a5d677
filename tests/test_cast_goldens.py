"""The reference's own unit-test vectors for the typesystem casts (tests/golden/cast_goldens.json, extracted from strictify_test.go,
restore_test.go and splitter_test.go by tests/golden/make_cast_goldens.py) against the oracle's restatement (oracle/cast_oracle.hpp)."""
import json
import os

import pytest

from transferia_b200 import abi

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cast_goldens.json"), encoding="utf-8"))
STRICT_GO = {"int8": ("int8",), "int16": ("int16",), "int32": ("int32",), "int64": ("int64",), "uint8": ("uint8",), "uint16": ("uint16",), "uint32": ("uint32",),
             "uint64": ("uint64",), "float": ("float32",), "double": ("json.Number",), "utf8": ("string",), "string": ("[]byte",), "boolean": ("bool",),
             "date": ("time.Time",), "datetime": ("time.Time",), "timestamp": ("time.Time",), "interval": ("time.Duration",)}      # type_checkers.go:39-84


def test_goldens_cover_the_reference_tests():
    assert len(G["strictify"]) == 53 and not G["strictify_skipped"]
    assert {c["test"] for c in G["strictify"]} >= {"TestStrictifyIntegerTypesPositive", "TestStrictifyIntegerTypesNegativeGreaterThanUpperBound",
                                                   "TestStrictifyFloatTypesPositive", "TestStrictifyTemporalTypesNegativeWrongType", "TestStrictifyBoolTypePositive"}
    assert len(G["restore"]) >= 40 and len(G["splitter"]) == 4


def test_strictify_matches_strictify_test_go(po):
    """strictify_test.go:54-684: positive items strictify without an error into the strict Go type of the column (type_checkers.go);
    negative items fail (StrictifyError) on at least one column — Strictify reports the first failing column."""
    for case in G["strictify"]:
        rcs = []
        for cell in case["values"]:
            rc, out = po.strictify_value(cell["value"]["go"], cell["value"]["v"], abi.YT_NAME_TO_TF[cell["type"]])
            rcs.append(rc)
            if rc == 0 and cell["value"]["go"] != "nil":
                assert out["go"] in STRICT_GO[cell["type"]], (case["test"], case["item"], cell, out)
        if case["expect"] == "ok":
            assert all(rc == 0 for rc in rcs), (case["test"], case["item"], rcs)
        else:
            assert any(rc in (1, 2) for rc in rcs) and all(rc in (0, 1, 2) for rc in rcs), (case["test"], case["item"], rcs)
    # the two range tests fail with StrictifyRangeError, not with a cast error, wherever the value is a Go integer (strictify.go:159-181)
    for case in G["strictify"]:
        if case["test"] in ("TestStrictifyIntegerTypesNegativeGreaterThanUpperBound", "TestStrictifyIntegerTypesNegativeLessThanLowerBound"):
            bad = [(c, po.strictify_value(c["value"]["go"], c["value"]["v"], abi.YT_NAME_TO_TF[c["type"]])[0]) for c in case["values"]]
            bad = [(c, rc) for c, rc in bad if rc]
            assert len(bad) == 1
            assert bad[0][1] == (1 if bad[0][0]["value"]["go"] == "string" else 2), bad


def _same(a, b):
    if a["go"] != b["go"]:
        return False
    if a["go"] in ("float32", "float64"):
        return float(a["v"]) == float(b["v"])
    if a["go"] == "map":
        return json.loads(a["v"]) == json.loads(b["v"])
    return a["v"] == b["v"]


def test_restore_matches_restore_test_go(po):
    """restore_test.go:14-135, every assertion the extractor could evaluate (the rest are listed in the fixture)."""
    for c in G["restore"]:
        rc, out = po.restore_value(c["in"]["go"], c["in"]["v"], c["type"])
        assert rc == 0, (c, rc)
        assert _same(out, c["want"]), (c, out)


def test_splitter_matches_splitter_test_go(po):
    for c in G["splitter"]:
        rows, rest = po.csv_split_rows(c["input"].encode())
        assert [r.decode() for r in rows] == c["rows"] and rest.decode() == c["eof_rest"], c
