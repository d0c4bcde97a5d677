"""Seeded synthetic ClickBench-`hits`-shaped batches in the C-ABI columnar layout.

The schema is the first 99 columns of the reference's fixture
(pkg/providers/postgres/testdata/hits_data.json: parse_schema), committed as
tests/golden/hits_schema.json by tests/golden/make_hits_schema.py.  Value generators:

  ints        column-specific range taken from the fixture cell magnitude: |v| <= 1 -> {0,1};
              int64 and id-like int32 (|v| >= 1e6) -> uniform over the full magnitude (incompressible);
              otherwise a 64-value palette drawn from [0, 2|v|] (sign kept), picked Zipf-like, which
              is how low-cardinality ClickBench dimensions (os, resolution, region ...) behave
  timestamps  in [2013-07-01, 2013-07-31) at second resolution, ascending within a batch (snapshot
              order); eventdate = day of eventtime; the other two are eventtime +- 1 day
  strings     per-column dictionary of 4096 base strings (length ~ geometric, mean = fixture cell
              length, 64-symbol alphabet with ~10 % two-byte Cyrillic), rows draw Zipf-like from it
              and half of the non-empty rows get a short random tail, so LZ4 sees realistic
              repetition; columns whose fixture cell is empty are 85 % empty, others 5 % empty
  sessions    rows arrive in visitor sessions (geometric, mean 8 rows): flag / palette ints and the
              string dimensions other than title/url/referer/searchphrase/params/originalurl keep one
              value per session, as consecutive hits of one visitor do
  nulls       none: every fixture column is `required: true`

Everything here is host-side numpy; the engine never sees this module.
"""
from __future__ import annotations

import base64
import json
import os
from typing import List, Tuple

import numpy as np

from . import abi

SEED = 0x7F4A7C15
_SCHEMA_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hits_schema.json")

T0 = 1372636800   # 2013-07-01T00:00:00Z
T1 = 1375228800   # 2013-07-31T00:00:00Z

_ALPHA = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789/.", dtype=np.uint8)


def hits_schema() -> List[dict]:
    cols = json.load(open(_SCHEMA_PATH))["columns"]
    out = []
    for c in cols:
        d = {k: v for k, v in c.items() if k != "fixture_cell_b64"}
        d["_fixture"] = base64.b64decode(c["fixture_cell_b64"])
        out.append(d)
    return out


def schema_json(schema: List[dict]) -> str:
    """ColSchema JSON with the reference's tags (col_schema.go:14-29)."""
    return json.dumps([{k: v for k, v in c.items() if not k.startswith("_")} for c in schema])


def _make_dict(rng: np.random.Generator, mean_len: float, n_entries: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    p = 1.0 / max(mean_len, 1.0)
    lens = np.minimum(rng.geometric(p, n_entries), 120).astype(np.int64)
    offs = np.zeros(n_entries + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    heap = _ALPHA[rng.integers(0, 64, total)]
    # ~10 % of positions become the lead of a 2-byte Cyrillic sequence (0xD0 0x90..0xBF)
    lead = np.nonzero(rng.random(total) < 0.05)[0]
    lead = lead[lead + 1 < total]
    # keep pairs inside one entry
    ent = np.searchsorted(offs, lead, side="right") - 1
    lead = lead[lead + 1 < offs[ent + 1]]
    heap = heap.copy()
    heap[lead] = 0xD0
    heap[lead + 1] = (0x90 + rng.integers(0, 48, lead.size)).astype(np.uint8)
    return offs, heap


def _gen_string_column(rng: np.random.Generator, n: int, mean_len: float, empty_frac: float, prefix: bytes = b"", sess=None):
    d_off, d_heap = _make_dict(rng, max(mean_len - len(prefix), 1.0))
    if prefix:   # e.g. URLs all start with http://
        nent = d_off.size - 1
        lens = np.diff(d_off) + len(prefix)
        new_off = np.zeros(nent + 1, dtype=np.int64); np.cumsum(lens, out=new_off[1:])
        new_heap = np.empty(int(new_off[-1]), dtype=np.uint8)
        pidx = (new_off[:-1, None] + np.arange(len(prefix))[None, :]).ravel()
        new_heap[pidx] = np.tile(np.frombuffer(prefix, dtype=np.uint8), nent)
        mask = np.ones(new_heap.size, dtype=bool); mask[pidx] = False
        new_heap[mask] = d_heap
        d_off, d_heap = new_off, new_heap
    nent = d_off.size - 1
    # Zipf-like pick: squared uniform concentrates on low indexes
    if sess is None:
        idx = (rng.random(n) ** 2.5 * nent).astype(np.int64)
        empty = rng.random(n) < empty_frac
        tail_len = np.where(rng.random(n) < 0.5, rng.integers(1, 9, n), 0)
    else:
        ns = int(sess[-1]) + 1
        idx = (rng.random(ns) ** 2.5 * nent).astype(np.int64)[sess]
        empty = (rng.random(ns) < empty_frac)[sess]
        tail_len = np.zeros(n, dtype=np.int64)
    base_len = (d_off[idx + 1] - d_off[idx])
    base_len = np.where(empty, 0, base_len); tail_len = np.where(empty, 0, tail_len)
    lens = base_len + tail_len
    offs = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    assert total < 2 ** 32
    heap = np.empty(total, dtype=np.uint8)
    # dictionary part
    rep_rows = np.repeat(np.arange(n), base_len)
    pos_in = np.arange(int(base_len.sum())) - np.repeat(np.cumsum(base_len) - base_len, base_len)
    heap[offs[rep_rows] + pos_in] = d_heap[d_off[idx[rep_rows]] + pos_in]
    # random tail
    tr = np.repeat(np.arange(n), tail_len)
    tpos = np.arange(int(tail_len.sum())) - np.repeat(np.cumsum(tail_len) - tail_len, tail_len)
    heap[offs[tr] + base_len[tr] + tpos] = _ALPHA[rng.integers(0, 64, tr.size)]
    return offs.astype(np.uint32), heap


def make_hits_batch(nrows: int, seed: int = SEED, schema: List[dict] | None = None) -> Tuple[abi.Batch, List[dict]]:
    schema = schema or hits_schema()
    rng = np.random.default_rng(seed)
    cols: List[abi.Column] = []
    eventtime = np.sort(rng.integers(T0, T1, nrows, dtype=np.int64))
    sess = np.cumsum(rng.random(nrows) < 0.125).astype(np.int64)
    nsess = int(sess[-1]) + 1 if nrows else 0
    for c in schema:
        t = abi.YT_NAME_TO_TF[c["type"]]
        fx = c["_fixture"]
        if t in (abi.TF_INT16, abi.TF_INT32, abi.TF_INT64):
            v = int(fx.decode() or "0")
            dt = abi.FIXED_DTYPE[t]
            info = np.iinfo(dt)
            if t == abi.TF_INT64:
                arr = rng.integers(info.min, info.max, nrows, dtype=np.int64, endpoint=True)
            elif abs(v) <= 1:
                arr = (rng.random(nsess) < 0.1)[sess].astype(dt)
            elif abs(v) >= 1_000_000:
                hi = min(2 * abs(v), info.max)
                a = rng.integers(0, hi + 1, nrows, dtype=np.int64)
                if v < 0:
                    a = -a
                arr = a.astype(dt)
            else:
                hi = min(2 * abs(v), info.max)
                palette = rng.integers(0, hi + 1, 64, dtype=np.int64)
                a = palette[(rng.random(nsess) ** 3 * 64).astype(np.int64)][sess]
                if v < 0:
                    a = -a
                arr = a.astype(dt)
            cols.append(abi.Column(t, values=arr))
        elif t == abi.TF_TIMESTAMP:
            if c["name"] == "eventtime":
                arr = eventtime.copy()
            else:
                arr = eventtime + rng.integers(-86400, 86400, nrows)
            cols.append(abi.Column(t, values=arr.astype(np.int64)))
        elif t == abi.TF_DATE:
            cols.append(abi.Column(t, values=(eventtime // 86400 * 86400).astype(np.int64)))
        elif t in (abi.TF_UTF8, abi.TF_ANY):
            mean_len = float(len(fx))
            empty_frac = 0.85 if len(fx) == 0 else 0.05
            if len(fx) == 0:
                mean_len = 12.0
            prefix = b"http://" if c["name"] in ("url", "referer", "originalurl") else b""
            per_row = c["name"] in ("title", "url", "referer", "searchphrase", "params", "originalurl")
            offs, heap = _gen_string_column(rng, nrows, mean_len, empty_frac, prefix, None if per_row else sess)
            aux = None
            if t == abi.TF_ANY:
                # pg "char" arrives as a Go string inside an `any` column (fixture: hitcolor = "5")
                aux = np.ones(nrows, dtype=np.uint8)
            cols.append(abi.Column(t, offsets=offs, heap=heap, aux=aux))
        else:
            raise NotImplementedError(c["type"])
    return abi.Batch(nrows, cols), schema


def counterid_threshold(batch: abi.Batch, schema: List[dict], keep_frac: float = 0.5) -> int:
    i = [c["name"] for c in schema].index("counterid")
    return int(np.quantile(batch.columns[i].values, 1.0 - keep_frac))


def headline_threshold(batch: abi.Batch, schema: List[dict], target: float = 0.28) -> int:
    """The watchid threshold K for which `watchid > K AND url ~ '://'` keeps `target` of the rows of THIS batch. Every rank of a
    weak-scaling run pushes its own seeded batch, and the work per GPU is only the same if the selectivity is: watchid is spread over
    the whole int64 range, so the quantile hits the target on any batch (counterid has 64 distinct values, one of them on 23 % of the rows)."""
    names = [c["name"] for c in schema]
    wid = np.asarray(batch.columns[names.index("watchid")].values)
    u = batch.columns[names.index("url")]
    url_ok = np.diff(np.asarray(u.offsets).astype(np.int64)) > 0          # every non-empty generated URL starts with "http://"
    if url_ok.sum() == 0:
        return int(wid.max())
    frac = min(1.0, target / max(url_ok.mean(), 1e-9))
    return int(np.quantile(wid[url_ok], 1.0 - frac, method="lower"))


def headline_transformers_watchid(k: int) -> List[dict]:
    """BASELINE.json configs[2]: cast + filter_rows (1 int term AND 1 string `~` term), the int term on watchid (see headline_threshold)."""
    return [{"filter_rows": {"tables": {"includeTables": ["^public\\.hits$"]},
                             "filter": f"watchid > {k} AND url ~ '://'"}}]


def headline_transformers(k: int) -> List[dict]:
    """BASELINE.json configs[2]: cast + filter_rows (1 int term AND 1 string `~` term)."""
    return [{"filter_rows": {"tables": {"includeTables": ["^public\\.hits$"]},
                             "filter": f"counterid > {k} AND url ~ '://'"}}]


# ----------------------------------------------------------------------------------------------- JSON lines (BASELINE configs[1])
JSON_FIELDS = [
    {"name": "id", "type": "int64", "key": True}, {"name": "ts", "type": "datetime"}, {"name": "user", "type": "utf8"}, {"name": "url", "type": "utf8"},
    {"name": "score", "type": "double"}, {"name": "ok", "type": "boolean"}, {"name": "cnt", "type": "uint32"}, {"name": "tags", "type": "any"},
]


def make_json_lines(nlines: int, seed: int = SEED + 1) -> Tuple[bytes, List[dict]]:
    """~160-byte JSON lines with 8 declared fields (plus one undeclared key every 16th line, a
    null every 11th, an escaped string every 13th), newline-terminated, as a queue producer would write them."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, 1 << 53, nlines); ts = rng.integers(1_372_636_800, 1_375_315_200, nlines)
    users = [f"user_{int(x):06d}" for x in rng.integers(0, 200_000, nlines)]
    hosts = ["example.com", "yandex.ru", "news.site.org", "shop.example.net", "пример.рф"]
    hsel = rng.integers(0, len(hosts), nlines); paths = rng.integers(0, 1 << 30, nlines)
    score = np.round(rng.random(nlines) * 1000, 6); okv = rng.random(nlines) < 0.5; cnt = rng.integers(0, 100_000, nlines)
    tagsel = rng.integers(0, 4, nlines)
    tags = ['["a","b"]', '{"k":1,"j":[1,2]}', '[]', '{"z":"q","a":null}']
    out = []
    for i in range(nlines):
        u = users[i] if i % 13 else "us\\\"er\\n\\u00e9_" + users[i]
        t = "null" if i % 11 == 0 else tags[tagsel[i]]
        extra = ',"trace":"%08x"' % paths[i] if i % 16 == 0 else ""
        out.append('{"id":%d,"ts":%d,"user":"%s","url":"https://%s/p/%d?ref=%d","score":%r,"ok":%s,"cnt":%d,"tags":%s%s}'
                   % (ids[i], ts[i], u, hosts[hsel[i]], paths[i], cnt[i], float(score[i]), "true" if okv[i] else "false", cnt[i], t, extra))
    return ("\n".join(out) + "\n").encode("utf-8"), [dict(f) for f in JSON_FIELDS]


# ----------------------------------------------------------------------------------------------------------- BASELINE configs[3]
DBZ_FIELDS = [("id", "int64", False), ("counter", "int32", True), ("region", "int32", True), ("flags", "int16", True), ("dur", "int64", True),
              ("price", "double", True), ("ok", "boolean", True), ("url", "string", True), ("title", "string", True), ("referer", "string", True),
              ("ts", "int64", True), ("small", "int8", True)]


def debezium_schema_text() -> str:
    """Kafka Connect schema of the 12-field CDC table (the `schema` member of a Debezium message / what the schema registry holds)."""
    fl = ",".join('{"type":"%s","optional":%s,"field":"%s"}' % (t, "true" if opt else "false", n) for n, t, opt in DBZ_FIELDS)
    return ('{"type":"struct","fields":[{"type":"struct","fields":[%s],"optional":true,"field":"before"},'
            '{"type":"struct","fields":[%s],"optional":true,"field":"after"}]}' % (fl, fl))


def make_debezium_messages(nmsgs: int, seed: int = SEED + 3, schema_id: int = 7, dml_frac: float = 0.02):
    """Debezium CDC envelopes with a 12-field payload, schema-registry framed (0x00 | u32be schema id | payload JSON,
    pkg/parsers/registry/debezium/engine/parser.go:42-50) so that the schema text travels once, not per message. Mostly inserts / snapshot
    reads (op c / r) plus a few updates and deletes, which filter_rows rejects per row (filter_rows.go:103-107).
    Returns (bytes, message end offsets, schema_text, (namespace, table))."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, 1 << 60, nmsgs); counter = rng.integers(0, 100_000, nmsgs); region = rng.integers(0, 300, nmsgs)
    flags = rng.integers(-3, 3, nmsgs); dur = rng.integers(0, 1 << 40, nmsgs); price = np.round(rng.random(nmsgs) * 10_000, 2)
    ok = rng.random(nmsgs) < 0.5; ts = rng.integers(1_372_636_800_000, 1_375_315_200_000, nmsgs); small = rng.integers(-128, 128, nmsgs)
    hosts = ["example.com", "shop.example.net", "news.site.org", "yandex.ru"]
    hsel = rng.integers(0, len(hosts), nmsgs); path = rng.integers(0, 1 << 30, nmsgs)
    opr = rng.random(nmsgs); lsn = 1_000_000 + np.arange(nmsgs) * 8
    frame = b"\x00" + int(schema_id).to_bytes(4, "big")
    out = []; ends = np.zeros(nmsgs, dtype=np.uint64); pos = 0
    for i in range(nmsgs):
        title = "Title %d" % path[i] if i % 17 else "T\\\"quoted\\\" \\u00e9 %d" % path[i]
        ref = "null" if i % 5 == 0 else '"https://%s/"' % hosts[(hsel[i] + 1) % len(hosts)]
        row = ('{"id":%d,"counter":%d,"region":%d,"flags":%d,"dur":%d,"price":%r,"ok":%s,"url":"https://%s/p/%d","title":"%s","referer":%s,"ts":%d,"small":%d}'
               % (ids[i], counter[i], region[i], flags[i], dur[i], float(price[i]), "true" if ok[i] else "false", hosts[hsel[i]], path[i], title, ref, ts[i], small[i]))
        if opr[i] < dml_frac / 2: op, before, after = "d", row, "null"
        elif opr[i] < dml_frac: op, before, after = "u", row, row
        else: op, before, after = ("r" if i % 3 == 0 else "c"), "null", row
        m = frame + ('{"before":%s,"after":%s,"source":{"version":"1.9","connector":"postgresql","name":"pg","ts_ms":%d,"snapshot":"false","db":"db","schema":"public",'
                     '"table":"events","txId":%d,"lsn":%d},"op":"%s","ts_ms":%d}' % (before, after, ts[i], 500 + i // 10, lsn[i], op, ts[i] + 3)).encode()
        out.append(m); pos += len(m); ends[i] = pos
    return b"".join(out), ends, debezium_schema_text(), ("public", "events")


def debezium_transformers(region_min: int = 150) -> List[dict]:
    """config #4's chain: SQL-predicate filter (one int term AND one substring term); the typesystem cast is the sink's."""
    return [{"filter_rows": {"filter": "region > %d AND url ~ 'example'" % region_min}}]


def render_hits_csv(batch: abi.Batch, schema: List[dict], seed: int = SEED + 5) -> bytes:
    """config #5's input: a hits-shaped batch as the CSV text a producer would write (`,` delimiter, `"` quotes doubled inside quoted cells,
    timestamps as "YYYY-MM-DD hh:mm:ss"), one row per line."""
    import datetime as dt
    rng = np.random.default_rng(seed)
    cols = []
    for c in batch.columns:
        if c.type in abi.VAR_TYPES:
            quote = rng.random(batch.nrows) < 0.1; vals = []
            for r in range(batch.nrows):
                s = bytes(c.heap[c.offsets[r]:c.offsets[r + 1]])
                if quote[r] or b"," in s or b'"' in s:
                    s = b'"' + s.replace(b'"', b'""') + b'"'
                vals.append(s)
            cols.append(vals)
        elif c.type == abi.TF_TIMESTAMP:
            cols.append([dt.datetime.fromtimestamp(int(v), dt.timezone.utc).strftime("%Y-%m-%d %H:%M:%S").encode() for v in c.values])
        elif c.type == abi.TF_DATE:
            cols.append([dt.datetime.fromtimestamp(int(v), dt.timezone.utc).strftime("%Y-%m-%d").encode() for v in c.values])
        else:
            cols.append([str(int(v)).encode() for v in c.values])
    return b"\n".join(b",".join(col[r] for col in cols) for r in range(batch.nrows)) + b"\n"
