"""GPU debug helper: run the JSON value matrix per field and print every cell / error that differs from the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from transferia_b200 import abi, engine
from oracle import pyoracle as po
import test_json_parser as T


def cellv(b, c, r):
    col = b.columns[c]
    ok = True if col.validity is None else bool((col.validity[r >> 3] >> (r & 7)) & 1)
    if col.type in abi.VAR_TYPES:
        v = bytes(col.heap[col.offsets[r]:col.offsets[r + 1]]); tag = int(col.aux[r]) if col.aux is not None else None
        return (ok, v, tag)
    return (ok, col.values[r:r + 1].tobytes().hex(), int(col.aux[r]) if col.aux is not None else None)


def diff(eng, text, fields, opts, msgs=None, label=""):
    schema = engine.json_result_schema(fields, opts)
    pid = eng.plan("db", "t", schema, [])
    got, gerr, gl = eng.parse_json(pid, text, opts, msgs)
    ref, rerr, rl = po.json_parse(text, fields, opts, msgs)
    lines = [ln for ln in text.split(b"\n") if ln.rstrip(b"\r")]
    ge = {r: (c, t) for r, c, t in gerr}; re_ = {r: (c, t) for r, c, t in rerr}
    bad = 0
    gi = ri = 0
    for k, ln in enumerate(lines) if msgs is None else []:
        g = ge.get(k); r = re_.get(k)
        if g != r:
            bad += 1; print(f"[{label}] line {k} {ln[:90]!r}: device err {g} oracle err {r}")
        if g is None and r is None:
            for c in range(len(schema)):
                a, b = cellv(got, c, gi), cellv(ref, c, ri)
                if a != b:
                    bad += 1; print(f"[{label}] line {k} {ln[:90]!r} col {schema[c]['name']}: device {a} oracle {b}")
        if g is None: gi += 1
        if r is None: ri += 1
    if gl != rl: print(f"[{label}] lines {gl} vs {rl}")
    return bad


eng = engine.Engine(0)
tot = 0
for f in T.ALL_FIELDS:
    for opts in ({}, {"use_numbers_in_any": True, "unpack_bytes_base64": True}, {"null_keys_allowed": True}):
        for req in (False, True):
            fields = [dict(f, required=req), {"name": "other", "type": "int32"}]
            text = T._lines_for(T.QUIRK_VALUES, [f]) + b'{"other":5}\n{"other":"x","%s":1}\n' % f["name"].encode()
            tot += diff(eng, text, fields, opts, label=f"{f['name']} {opts} req={req}")
print("TOTAL DIFFS", tot)
