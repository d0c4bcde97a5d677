"""Generic JSON parser: the oracle against the reference's canon data on CPU, the device parser against
the oracle on GPU (lines -> typed columns -> transformer chain -> sink bytes without leaving HBM)."""
import base64
import json
import os

import numpy as np
import pytest

from transferia_b200 import abi

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "json_parser_goldens.json")))

NUM_FIELDS = [{"name": "id", "type": "int8"}, {"name": "number_field", "type": "int64"}, {"name": "float_field", "type": "double"},
              {"name": "obj_field", "type": "any"}, {"name": "array_field", "type": "any"}]            # parser_test.go:132-153
B64_FIELDS = [{"name": "id", "type": "int8"}, {"name": "stringVal", "type": "utf8"}, {"name": "bytesVal", "type": "string"}]   # :234-247


def cell(batch, c, r):
    col = batch.columns[c]
    if col.validity is not None and not (col.validity[r >> 3] >> (r & 7)) & 1:
        return None
    if col.type in abi.VAR_TYPES:
        raw = bytes(col.heap[col.offsets[r]:col.offsets[r + 1]])
        if col.type == abi.TF_ANY:
            return raw.decode() if col.aux[r] == 1 else json.loads(raw)
        return raw
    return col.values[r].item()


def test_number_types_canon(po):
    """TestParserNumberTypes (parser_test.go:129-231): one message per line, both UseNumbersInAny modes."""
    text = G["inputs"]["parser_numbers_test.jsonl"].encode()
    for mode, use in (("UseNumbersFalse", False), ("UseNumbersTrue", True)):
        want = G["canon"]["TestParserNumberTypes"][mode]
        b, errs, lines = po.json_parse(text, NUM_FIELDS, {"use_numbers_in_any": use})
        assert not errs and lines == b.nrows == len(want)
        for r, it in enumerate(want):
            assert it["types"] == ["int8", "int64", "double", "any", "any"]
            got = [cell(b, c, r) for c in range(5)]
            assert got == it["columnvalues"], (mode, got)
    # the spelling Go's encoding/json gives the `any` cells (sorted keys, float64 'f'/'e' switch-over, json.Number verbatim)
    b, _, _ = po.json_parse(text, NUM_FIELDS, {})
    assert bytes(b.columns[3].heap) == b'{"int_field":123123123123123330}'
    assert bytes(b.columns[4].heap) == b"[0,1,2,4,12313.12241632513,-123123.13117532,12345678987654320,-12345678987654320,100000,1e-7]"
    b, _, _ = po.json_parse(text, NUM_FIELDS, {"use_numbers_in_any": True})
    assert bytes(b.columns[4].heap) == b"[0,1,2,4,12313.12241632513,-123123.13117532,12345678987654321,-12345678987654321,1e5,0.0000001]"


def test_base64_canon(po):
    """TestBase64Unpack (parser_test.go:233-276): `string` (bytes) cells are base64-decoded, utf8 cells are unescaped."""
    text = G["inputs"]["parse_base64_packed.jsonl"].encode()
    want = G["canon"]["TestBase64Unpack"]
    b, errs, lines = po.json_parse(text, B64_FIELDS, {"unpack_bytes_base64": True})
    assert not errs and b.nrows == len(want) == 2
    for r, it in enumerate(want):
        assert cell(b, 0, r) == it["columnvalues"][0]
        assert cell(b, 1, r).decode() == it["columnvalues"][1]
        assert cell(b, 2, r) == base64.b64decode(it["columnvalues"][2])       # the canon file prints []byte as base64
    # without the option the cell keeps the text
    b, _, _ = po.json_parse(text, B64_FIELDS, {})
    assert cell(b, 2, 0) == b"dGVzdA=="
