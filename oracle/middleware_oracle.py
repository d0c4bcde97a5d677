"""CPU restatement (TEST INFRASTRUCTURE — never imported by the product) of the reference's sink pipeline below the user transformers,
item by item over Python lists, used to check tfgpu_sink_push's sequencing and counters:

    transformation.Push      pkg/transformer/transformation.go:122-158 (SplitByTableID pkg/abstract/changeitem/utils.go:130-136; per-table do()
                             :236-282; errors first :152-157; errorChangeItems :206-234)
    NonRowSeparator.Push     pkg/middlewares/nonrow_separator.go:29-55
    filter.Push              pkg/middlewares/filter.go:60-77 with ExcludeSystemTables (ChangeItem.IsSystemTable: Table in the registered set,
                             pkg/abstract/changeitem/system_table.go:26-30)
    statistician.Push        pkg/middlewares/statistician.go:55-68 -> WrapperStats.Log pkg/stats/sink_wrapper.go:63-78, batchStats
                             pkg/stats/sink_wrapper_util.go:10-50

Parity pinned by: nothing the reference holds as bytes for these middlewares (they have no golden files); the restatement follows the
cited lines one to one and the reference's own unit expectations for NonRowSeparator (pkg/middlewares/nonrow_separator_test.go) are
re-stated in tests/test_sink_push.py.

`apply_row_transformers(table_index, items) -> (kept_items, error_items)` is supplied by the test (the oracle's plan over the table's row
events); non-row items pass through every transformer except skip_events (drops the listed kinds) and rename_tables."""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

ROW_KINDS = (0, 1, 2)
KIND_SYNCHRONIZE = 24
KIND_NAME = {0: "insert", 1: "update", 2: "delete", 16: "init_sharded_table_load", 17: "init_load_table", 18: "done_load_table",
             19: "done_sharded_table_load", 20: "drop_table", 21: "truncate", 22: "DDL", 23: "pg:DDL", 24: ""}


def non_row_separator(items: List[dict]) -> List[List[dict]]:
    """nonrow_separator.go:29-55: the downstream Push calls for one incoming batch."""
    pushes, start, end = [], 0, 0
    for i, it in enumerate(items):
        if it["kind"] in ROW_KINDS:
            end += 1; continue
        if end > start:
            pushes.append(items[start:end])
        pushes.append(items[i:i + 1])
        start = end = i + 1
    if end > start:
        pushes.append(items[start:end])
    return pushes


def filter_push(items: List[dict], system_tables: Sequence[str], stats: dict) -> List[dict]:
    """filter.go:60-77: drop items of system tables; an empty batch is not forwarded (the caller checks)."""
    out = []
    for it in items:
        if it["out_table"] in system_tables:
            stats["filter_dropped"] += 1; continue
        out.append(it)
    return out


def statistician(items: List[dict], stats: dict) -> None:
    """sink_wrapper.go:63-78 after a successful downstream Push + batchStats sink_wrapper_util.go:10-50 (the variable the reference calls
    `oldestTime` ends up holding the LATEST commit time: `if oldestTime.Before(eventTime) { oldestTime = eventTime }`)."""
    stats["downstream_pushes"] += 1
    stats["change_items_pushed"] += len(items)
    for it in items:
        if not (it["kind"] in ROW_KINDS or it["kind"] == KIND_SYNCHRONIZE):
            continue
        if it["kind"] in ROW_KINDS:
            stats["row_events_pushed"] += 1
        stats["inflight_bytes"] += it.get("size", 0)
        ct = it.get("commit_time", 0)
        if not ct:
            stats["without_commit_time"] += 1; continue
        stats["max_commit_time"] = max(stats["max_commit_time"], ct)
        stats["min_commit_time"] = ct if not stats["min_commit_time"] else min(stats["min_commit_time"], ct)


def sink_push(items: List[dict], tables: List[Tuple[str, str]], skip_events: Sequence[Tuple[Callable[[str, str], bool], Sequence[str]]],
              renames: Dict[Tuple[str, str], Tuple[str, str]], system_tables: Sequence[str],
              apply_row_transformers: Callable[[int, List[dict]], Tuple[List[dict], List[dict]]], errors_to_sink: bool, stats: dict) -> List[dict]:
    """One Sinker.Push through the pipeline. items: dicts with kind, table (index), index, commit_time, size. Returns the downstream pushes as
    {"type": "rows" | "item" | "errors", "table": k, "items": [input indexes]}. Tables are visited in order of first appearance and a push
    never spans two tables / schemas (the product's granularity; the reference may glue the last row run of one table to the first of the
    next — totals are the same)."""
    stats["pushes"] += 1; stats["max_commit_time"] = stats["min_commit_time"] = 0
    groups: Dict[Tuple[str, str], List[dict]] = {}
    for it in items:                                                       # SplitByTableID
        groups.setdefault(tables[it["table"]], []).append(it)
    out: List[dict] = []

    def deliver(kind: str, table: int, its: List[dict]):
        out.append({"type": kind, "table": table, "items": [x["index"] for x in its]})
        statistician(its, stats)

    for tid, its in groups.items():
        # transformation.do over the table's items: row events through the plan, the others through skip_events / rename only
        seq: List = []                                                     # ("run", table, [items]) | ("item", item)
        for it in its:
            if it["kind"] in ROW_KINDS:
                if seq and seq[-1][0] == "run" and seq[-1][1] == it["table"]:
                    seq[-1][2].append(it)
                else:
                    seq.append(("run", it["table"], [it]))
                continue
            if any(suit(*tid) and KIND_NAME.get(it["kind"], "\x01other") in evs for suit, evs in skip_events):
                stats["transform_dropped"] += 1; continue                  # gone before NonRowSeparator: the neighbouring row runs join
            it = dict(it); it["out_table"] = renames.get(tid, tid)[1]
            seq.append(("item", it))
        for ent in seq:
            if ent[0] == "item":
                kept = filter_push([ent[1]], system_tables, stats)
                if kept:
                    deliver("item", ent[1]["table"], kept)
                continue
            _, table, run = ent
            out_table = renames.get(tid, tid)[1]
            if out_table in system_tables:
                stats["filter_dropped"] += len(run); continue
            kept, errs = apply_row_transformers(table, run)
            stats["transform_dropped"] += len(run) - len(kept); stats["transform_errors"] += len(errs)
            if errs and errors_to_sink:
                deliver("errors", table, errs)
            if kept:
                deliver("rows", table, kept)
    return out


# ------------------------------------------------------------------ table_splitter (registry/table_splitter/table_splitter.go:36-58)
def _go_v_float(x: float, is32: bool) -> str:
    """fmt %v of a float: strconv 'g' with the shortest digits, exponent form when exp < -4 || exp >= 6 (strconv/ftoa.go, eprec = 6)."""
    import math
    import numpy as np
    if math.isnan(x): return "NaN"
    if math.isinf(x): return "+Inf" if x > 0 else "-Inf"
    r = np.format_float_scientific(np.float32(x) if is32 else np.float64(x), unique=True, trim="-")      # shortest digits, d.ddde+XX
    mant, _, ex = r.partition("e"); exp = int(ex)
    neg = mant.startswith("-"); mant = mant.lstrip("-"); digits = mant.replace(".", "")
    if exp < -4 or exp >= 6:
        out = digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + "e%s%02d" % ("-" if exp < 0 else "+", abs(exp))
    elif exp < 0:
        out = "0." + "0" * (-exp - 1) + digits
    else:
        out = digits + "0" * (exp + 1 - len(digits)) if len(digits) <= exp + 1 else digits[:exp + 1] + "." + digits[exp + 1:]
    return ("-" if neg else "") + out


def serialize_to_string(value, yt_type: str) -> str:
    """to_string.SerializeToString (registry/to_string/to_string.go:145-172) over the Go-typed (tag, value) pairs of transferia_b200.rows
    (tags: 0 nil, 1 bool, 2-5 int8..64, 6-9 uint8..64, 10 float32, 11 float64, 12 string, 13 []byte, 14 time.Time (sec, nsec), 16 json.Number)."""
    import datetime as dt
    tag, v = value
    if tag == 0: return "<nil>"
    if tag == 1: return "true" if v else "false"
    if 2 <= tag <= 9: return str(int(v))
    if tag == 10: return _go_v_float(v, True)
    if tag == 11: return _go_v_float(v, False)
    if tag in (12, 16): return v.decode("utf-8", "surrogateescape")
    if tag == 13 and yt_type == "string": return v.decode("utf-8", "surrogateescape")
    if tag == 14 and yt_type in ("date", "datetime", "timestamp"):
        import numpy as np
        text = str(np.datetime64(int(v[0]), "s"))                 # proleptic Gregorian, any year: "YYYY-MM-DDTHH:MM:SS" ("-1232-..", "14124-..")
        if yt_type == "date": return text.split("T")[0]
        frac = ("." + ("%09d" % v[1]).rstrip("0")) if v[1] else ""
        return text + frac + "Z"
    raise NotImplementedError((tag, yt_type))


def generate_table_name(original: str, columns, splitter: str, values_by_name: dict, types_by_name: dict) -> str:
    """GenerateTableName table_splitter.go:36-58: the current table name (when not empty), then the text of every listed column the schema
    knows (a column the item does not carry reads nil)."""
    parts = [original] if original else []
    for col in columns:
        if col in types_by_name:
            parts.append(serialize_to_string(values_by_name.get(col, (0, None)), types_by_name[col]))
    return (splitter or "/").join(parts)


# ------------------------------------------------------------------ updatable ClickHouse tables (pkg/providers/clickhouse/sink_table.go:397-432)
def updatable_args(kind: int, values, old_keys: dict, commit_time: int, ncols: int):
    """buildChangeItemArgs for one item of an updatable table, as the list of per-column values (None = nil) followed by
    (__data_transfer_commit_time, __data_transfer_delete_time): an insert keeps its values + (CommitTime, 0); a delete is built from OldKeys
    (buildDeleteKindArgs: the old value of every column OldKeys lists, nil for the others — fillRequiredColumn is false on current servers)
    + (CommitTime, CommitTime)."""
    if kind == 0:
        return list(values) + [commit_time, 0]
    if kind == 2:
        return [old_keys.get(c) for c in range(ncols)] + [commit_time, commit_time]
    raise NotImplementedError("updates go through Collapse / the toast lookup")
