"""An independent READER of the ClickHouse native Data block (test infrastructure): it knows nothing of how the oracle or the device lay a
block out — it follows the public native format at client revision 54454+ from the bytes alone (BlockInfo fields, column count, row count, per
column: name, type, custom-serialization byte, then the column data by TYPE: fixed-width little-endian values, LEB128-prefixed strings, a null
map of one byte per row in front of Nullable columns). What it reads back is compared with the input batch after the casts the reference applies
(columntypes.Restore: date / datetime clamps, DateTime64(6) micros, bool -> UInt8), so an encoder that disagrees with the format on widths,
order, null maps or string framing is caught by a second, structurally different piece of code."""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

FIXED = {"Int8": "<i1", "Int16": "<i2", "Int32": "<i4", "Int64": "<i8", "UInt8": "<u1", "UInt16": "<u2", "UInt32": "<u4", "UInt64": "<u8",
         "Float32": "<f4", "Float64": "<f8", "Date": "<u2", "DateTime": "<u4"}


def _uvarint(b: bytes, p: int) -> Tuple[int, int]:
    v = s = 0
    while True:
        x = b[p]; p += 1; v |= (x & 0x7f) << s; s += 7
        if not x & 0x80:
            return v, p


def _string(b: bytes, p: int) -> Tuple[bytes, int]:
    n, p = _uvarint(b, p)
    return b[p:p + n], p + n


def read_block(b: bytes) -> Tuple[List[Tuple[str, str]], int, Dict[str, list]]:
    """-> ([(name, type)], rows, {name: values}); Nullable columns hold None for nulls; raises on trailing bytes."""
    p = 0
    while True:                                               # BlockInfo: (field number, value)*, 0 ends it
        f, p = _uvarint(b, p)
        if f == 0: break
        if f == 1: p += 1                                     # is_overflows: UInt8
        elif f == 2: p += 4                                   # bucket_num: Int32
        else: raise ValueError(f"unknown BlockInfo field {f}")
    ncols, p = _uvarint(b, p); nrows, p = _uvarint(b, p)
    cols, data = [], {}
    for _ in range(ncols):
        name, p = _string(b, p); typ, p = _string(b, p)
        assert b[p] == 0, "custom serialization is not expected"; p += 1
        name, typ = name.decode(), typ.decode(); cols.append((name, typ))
        inner, nulls = typ, None
        if typ.startswith("Nullable("):
            inner = typ[9:-1]
            if nrows: nulls = np.frombuffer(b, np.uint8, nrows, p).astype(bool); p += nrows
        if inner == "String":
            vals = []
            for _r in range(nrows):
                s, p = _string(b, p); vals.append(s)
        else:
            dt = FIXED.get(inner) or ("<i8" if inner.startswith("DateTime64") else None)
            if dt is None: raise ValueError(f"type {inner} is not known to this reader")
            w = np.dtype(dt).itemsize
            vals = list(np.frombuffer(b, dt, nrows, p)); p += w * nrows
        if nulls is not None:
            vals = [None if nulls[r] else vals[r] for r in range(nrows)]
        data[name] = vals
    if p != len(b): raise ValueError(f"{len(b) - p} bytes behind the last column")
    return cols, nrows, data
