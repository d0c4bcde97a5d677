/*
 * tfgpu_sink.h — host side of the drop-in boundary: what stands between Sinker.Push([]abstract.ChangeItem) and the device calls of
 * tfgpu.h, and between the device's wire bytes and the destination's socket. Host-only C++ inside libtfgpu.so (no kernel is launched
 * by anything declared here except through the tfgpu_engine a sink owns); C-ABI, plain pointers and sizes.
 *
 *   tfgpu_ch_*        ClickHouse native TCP client writer: the Hello / Query / Data exchange clickhouse-go/v2 runs for
 *                     sql.Tx.Prepare + Exec + Commit (pkg/providers/clickhouse/sink_table.go:605-684) and for the streaming batch of
 *                     pkg/providers/clickhouse/async/streamer.go:64-265 (PrepareBatch / Append / Flush / Send), with the connection options
 *                     of pkg/providers/clickhouse/conn/connection.go:14-52 (database, user, password, LZ4 compression). Data packets carry
 *                     the compressed frames the device produced (TF_WIRE_CH_NATIVE_LZ4) without another copy or re-compression.
 *   tfgpu_rows_*      host transpose and its inverse: []ChangeItem in a flat row-major image (`tf_rows`, written by the Go
 *                     shim with plain appends, no cgo call per value) <-> the columnar tf_batch of tfgpu.h in pooled buffers,
 *                     pkg/abstract/changeitem/change_item.go:27-78, old_keys.go:3-7, kind.go:5-43.
 *   tfgpu_sink_*      the always-on middleware below the transformers and the Push of the shim as real code:
 *                     transformation (per-table plans) -> NonRowSeparator (pkg/middlewares/nonrow_separator.go:29-55) ->
 *                     Filter(ExcludeSystemTables) (pkg/middlewares/filter.go:60-77) -> Statistician counters
 *                     (pkg/middlewares/statistician.go:55-68, pkg/stats/sink_wrapper.go:56-104) -> destination.
 */
#ifndef TFGPU_SINK_H_
#define TFGPU_SINK_H_

#include "tfgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TF_E_RETRY_IO      3   /* retriable: socket read / write failed or timed out                                  */
#define TF_E_RETRY_SERVER  4   /* retriable: the ClickHouse server answered with an Exception packet (tfgpu_ch_last_error) */
#define TF_E_FATAL_PROTOCOL -5 /* fatal: the peer does not speak the native protocol / a block type the reader cannot skip */

/* ------------------------------------------------------------------ ClickHouse native client writer */
typedef struct tfgpu_ch_conn tfgpu_ch_conn;

/* clickhouse_go.Open over an already connected stream socket (the shim dials; TLS is out of scope): client Hello, server Hello, and the
 * quota-key addendum when the negotiated revision has it. The client speaks protocol revision 54460 (clickhouse-go/v2 v2.46.0,
 * DBMS_TCP_PROTOCOL_VERSION) and negotiates min(54460, server revision) like the driver.
 * opts_json: {"database":"default","user":"default","password":"","client_name":"transferia-tfgpu","compression":true,
 *             "read_timeout_ms":300000}   (conn/connection.go:38-51: LZ4 compression, 5 min read timeout).
 * The fd stays owned by the caller; tfgpu_ch_close only frees the handle. When the call fails *out still holds a handle (tfgpu_ch_last_error /
 * tfgpu_ch_exception_code say why): close it. */
int tfgpu_ch_open(int fd, const char* opts_json, tfgpu_ch_conn** out);
int tfgpu_ch_close(tfgpu_ch_conn* c);
const char* tfgpu_ch_last_error(const tfgpu_ch_conn* c);
/* {"name":..,"major":..,"minor":..,"patch":..,"revision":<negotiated>,"server_revision":..,"timezone":..,"display_name":..} */
const char* tfgpu_ch_server_info(const tfgpu_ch_conn* c);
/* ClickHouse error code of the last Exception packet (0 = none). */
int tfgpu_ch_exception_code(const tfgpu_ch_conn* c);

/* conn.PrepareBatch: Query packet (query id, client info, settings, stage Complete, compression flag, body) followed by the empty Data
 * block that ends the external tables, then reads packets until the server's sample block (TableColumns / Progress / Log / ProfileEvents
 * packets in between are consumed).  query: "INSERT INTO `db`.`t` (`a`,`b`) VALUES" (tfgpu_ch_insert_query builds the text the reference
 * builds); settings_json: {"insert_distributed_sync":"1",...} (model.InsertParams.AsQueryPart / ToQueryOption) or NULL. */
int tfgpu_ch_insert_begin(tfgpu_ch_conn* c, const char* query, const char* query_id, const char* settings_json);
/* The sample block's columns as JSON [{"name":"a","type":"Int32"},..]; valid until the next insert_begin. */
const char* tfgpu_ch_insert_columns(const tfgpu_ch_conn* c);
/* batch.Flush / Send for one block: a Data packet whose body is `wire` — the frame stream of TF_WIRE_CH_NATIVE_LZ4 (compression on) or
 * the raw block of TF_WIRE_CH_NATIVE (compression off) exactly as tfgpu_result_bytes returns it. Header and frames leave in one writev. */
int tfgpu_ch_insert_data(tfgpu_ch_conn* c, const uint8_t* wire, uint64_t len);
/* batch.Send's tail: the empty block that ends the INSERT, then reads until EndOfStream. written_* (may be NULL) receive the server's
 * Progress totals (written rows / bytes). A server Exception gives TF_E_RETRY_SERVER with its text in tfgpu_ch_last_error. */
int tfgpu_ch_insert_end(tfgpu_ch_conn* c, uint64_t* written_rows, uint64_t* written_bytes);
/* Counters since open: bytes written to / read from the socket, Data packets sent. */
int tfgpu_ch_stats(const tfgpu_ch_conn* c, uint64_t* bytes_out, uint64_t* bytes_in, uint64_t* data_packets);

/* doOperation's statement text (sink_table.go:633-660): INSERT INTO `db`.`table` (`c1`,`c2`[,`__data_transfer_commit_time`,
 * `__data_transfer_delete_time`]) VALUES — clickhouse-go cuts the statement at VALUES before it sends it. columns_json: the plan's result
 * column names as a JSON array of strings. Returns the byte count written (without NUL) or a negative code when cap is too small. */
int64_t tfgpu_ch_insert_query(const char* database, const char* table, const char* columns_json, int updateable, char* out, uint64_t cap);

/* ------------------------------------------------------------------ []ChangeItem in row form (host transpose) */
/* abstract.Kind (pkg/abstract/changeitem/kind.go:5-43). Row kinds keep tf_batch's TF_KIND_INSERT / UPDATE / DELETE. */
#define TF_KIND_INIT_SHARDED_TABLE_LOAD 16
#define TF_KIND_INIT_TABLE_LOAD         17
#define TF_KIND_DONE_TABLE_LOAD         18
#define TF_KIND_DONE_SHARDED_TABLE_LOAD 19
#define TF_KIND_DROP_TABLE              20
#define TF_KIND_TRUNCATE                21
#define TF_KIND_DDL                     22
#define TF_KIND_PG_DDL                  23
#define TF_KIND_SYNCHRONIZE             24   /* Kind("") */
#define TF_KIND_OTHER                   31   /* mongo:* / ch:* / es:* — opaque to this path, passed through */
#define TF_KIND_IS_ROW(k) ((k) <= TF_KIND_DELETE)   /* RowEventKinds: ChangeItem.IsRowEvent (change_item.go:286-288) */

/* One ColumnValues entry in the row image: a tag byte naming the Go dynamic type, then the payload (little-endian).
 * The shim writes these with plain appends while it walks []interface{} — one type switch per value, no cgo call. */
#define TF_V_NIL       0
#define TF_V_BOOL      1   /* 1 byte */
#define TF_V_INT8      2   /* natural width; Go `int` travels as TF_V_INT64 */
#define TF_V_INT16     3
#define TF_V_INT32     4
#define TF_V_INT64     5
#define TF_V_UINT8     6
#define TF_V_UINT16    7
#define TF_V_UINT32    8
#define TF_V_UINT64    9
#define TF_V_FLOAT32   10
#define TF_V_FLOAT64   11
#define TF_V_STRING    12  /* u32 length + bytes */
#define TF_V_BYTES     13  /* u32 length + bytes */
#define TF_V_TIME      14  /* time.Time: i64 Unix seconds + u32 nanoseconds (the UTC instant) */
#define TF_V_DURATION  15  /* i64 nanoseconds */
#define TF_V_JSONNUM   16  /* json.Number: u32 length + text */
#define TF_V_JSON      17  /* anything else (map, slice, ...): u32 length + its json.Marshal text */

/* What the items of one table share and never rewrite (change_item.go:44-52: ColumnNames / TableSchema are write-once). */
typedef struct tf_table {
    const char* schema;        /* ChangeItem.Schema (namespace) */
    const char* table;         /* ChangeItem.Table */
    const char* schema_json;   /* TableSchema.Columns() with the ColSchema JSON tags (col_schema.go:14-29); NULL for items without a schema (DDL, drop) */
} tf_table;

/* ChangeItem (change_item.go:27-78) without its values. */
#define TF_ITEM_SPARSE 1       /* ColumnNames is a subset of the schema (toasted update): every value is preceded by its u16 column index */
typedef struct tf_item {
    uint64_t lsn;              /* ChangeItem.LSN */
    uint64_t commit_time;      /* ChangeItem.CommitTime, ns */
    uint64_t size_read;        /* ChangeItem.Size.Read */
    uint64_t size_values;      /* ChangeItem.Size.Values (set by the Measurer, tfgpu_measure): what batchStats sums (sink_wrapper_util.go:24) */
    uint64_t values_off;       /* row kinds: offset of the ColumnValues image in tf_rows.values */
    uint64_t old_keys_off;     /* offset of the OldKeys image (u16 count, then {u16 column, value}*), or UINT64_MAX */
    uint32_t id;               /* ChangeItem.ID */
    uint32_t table;            /* index into tf_rows.tables */
    uint32_t n_values;         /* len(ColumnValues) */
    uint32_t txid_off, txid_len;   /* ChangeItem.TxID in tf_rows.strings */
    uint32_t part_off, part_len;   /* ChangeItem.PartID */
    int32_t  counter;          /* ChangeItem.Counter */
    uint8_t  kind;             /* TF_KIND_* */
    uint8_t  flags;            /* TF_ITEM_* */
    uint8_t  pad[2];
} tf_item;

typedef struct tf_rows {
    uint64_t        n_items;
    const tf_item*  items;
    uint32_t        n_tables;
    uint32_t        pad;
    const tf_table* tables;
    const uint8_t*  values;   uint64_t values_len;
    const uint8_t*  strings;  uint64_t strings_len;
} tf_rows;

/* Pooled columnar buffers (grow-only, pinned when a CUDA device is present so that tfgpu_push_* copies straight from them). */
typedef struct tfgpu_columnar tfgpu_columnar;
int tfgpu_columnar_create(tfgpu_columnar** out);
int tfgpu_columnar_destroy(tfgpu_columnar* p);
const char* tfgpu_columnar_last_error(const tfgpu_columnar* p);

/* The transpose: the row items `item_idx[0..n)` of `rows` (all of one table, row kinds, in this order; item_idx NULL = every row item of
 * table `table`) become one tf_batch in the pool's buffers, valid until the next call on this pool:
 *   - a column whose values all carry the column's canonical Go type (typesystem/values/type_checkers.go:39-84) gets the schema's type;
 *     ints / floats of another width give a LOOSE column (tf_col.type = what arrived) that the device strictifies (strictify.go:46-157);
 *     var-width cells shorter than 256 / 65536 bytes travel as TF_COL_LENS8 / TF_COL_LENS16 lengths;
 *   - kinds -> tf_batch.kinds; ID / LSN / CommitTime / TxID -> *meta; OldKeys -> *old (NULL when no item carries any).
 * Values a column cannot hold (a string in an int column, mixed signed / unsigned / float classes in one column) fail the call with
 * TF_E_FATAL_UNSUPPORTED: that batch stays on the Go path. threads <= 0: one per core (capped at 16). */
int tfgpu_rows_to_batch(tfgpu_columnar* pool, const tf_rows* rows, uint32_t table, const uint64_t* item_idx, uint64_t n, int threads,
                        const tf_batch** batch, const tf_row_meta** meta, const tf_old_keys** old);
/* The inverse, for tfgpu_push_columns results going back into []ChangeItem: every row of `b` as a ColumnValues image (the canonical Go
 * type of each column: int8..uint64, float32, float64, bool, string for utf8, []byte for string, time.Time, time.Duration, and for `any`
 * a Go string when the cell's tag says so, else its JSON text). row_off receives nrows + 1 offsets into out. Returns TF_E_FATAL_ARG with
 * *need set when cap is too small. */
int tfgpu_batch_to_rows(const tf_batch* b, uint8_t* out, uint64_t cap, uint64_t* row_off, uint64_t* need);

/* Rows keep[r] != 0 of a host batch, in order, in the pool's buffers (valid until the next call on the pool): the host half of
 * tfgpu_push_encode_selective. sel_out (optional) receives the input row of every output row. Columns whose buffers are NULL stay NULL. */
int tfgpu_batch_gather(tfgpu_columnar* pool, const tf_batch* in, const uint8_t* keep, int threads, const tf_batch** out, const uint32_t** sel_out);
/* The same with the selection at hand: sel = m ascending row indexes (tfgpu_push_encode_selective gets them from the device). */
int tfgpu_batch_gather_sel(tfgpu_columnar* pool, const tf_batch* in, const uint32_t* sel, uint64_t m, int threads, const tf_batch** out);

/* ------------------------------------------------------------------ Sinker.Push as one call */
/* What stands below the user's transformers in every transfer's sink pipeline (pkg/sink_factory/sink_factory.go:79-108), in the
 * reference's order, over items in row form:
 *   transformation.Push   pkg/transformer/transformation.go:122-158: items split by table (SplitByTableID), one plan per (table, schema)
 *                         cached like transformation.go:93-121, the plan's chain over the table's ROW events on the device; items of other
 *                         kinds go through untouched except for skip_events (drops the listed kinds) and rename_tables (renames them);
 *                         TransformerResult.Errors are pushed downstream first ("sink") or dropped ("devnull"), transformation.go:173-205
 *   NonRowSeparator       pkg/middlewares/nonrow_separator.go:29-55: runs of row events stay together, every other item travels alone
 *   Filter                pkg/middlewares/filter.go:60-77 with ExcludeSystemTables: items of registered system tables are dropped
 *   Statistician          pkg/middlewares/statistician.go:55-68 -> stats/sink_wrapper.go:63-78, sink_wrapper_util.go:10-50: counters per
 *                         downstream Push that succeeded
 *   destination           the row runs encoded by the device (cfg "wire_fmt") go to the ClickHouse writer (tfgpu_sink_set_clickhouse: one
 *                         INSERT per run, the statement of sink_table.go:633-660) or to the callback; everything else goes to the callback.
 * Tables are visited in order of first appearance (the reference ranges over a Go map: any order is legal there).
 * cfg_json: {"transformers":[..], "errors_output":"sink"|"devnull", "exclude_system_tables":true, "system_tables":["__consumer_keeper",..],
 *            "sink":{"type":"clickhouse"}, "database":"db", "updateable":false, "wire_fmt":2,
 *            "debezium":{<opts_json of tfgpu_emit_debezium>}}   (wire_fmt TF_WIRE_DEBEZIUM: the queue Debezium serializer
 *            pkg/serializer/queue/debezium_serializer.go:25-92 — key / value / tombstone messages of every row kind through
 *            tfgpu_emit_debezium_crud, with OldKeys and the `source` block's ID / LSN / CommitTime / TxID from the row form)
 * wire_fmt 0: row runs are handed on as columnar batches (ev.batch; with transformers the result of tfgpu_push_columns). `e` may be NULL
 * only when wire_fmt is 0 and the transformer list holds nothing but skip_events / rename_tables / table_splitter / regex_replace_transformer:
 * those act on kinds, table names and (regex_replace) the row image, which the host handles; everything else needs the device.
 * "updateable": true (model.ChSinkParams.IsUpdateable): the destination table carries `__data_transfer_commit_time` / `__data_transfer_delete_time`
 * (UInt64) behind its columns; buildChangeItemArgs / buildDeleteKindArgs (sink_table.go:397-432) run on the row image before the transpose — an
 * insert keeps its values + (CommitTime, 0), a delete is rebuilt from OldKeys (nil for the columns OldKeys does not list: insert_null_as_default
 * fills them on the server) + (CommitTime, CommitTime) — and the device encodes ordinary rows of the extended schema. Update items (Collapse,
 * toast lookup: sink_table.go:618-626) and transformer chains are refused in this mode.
 * table_splitter (pkg/transformer/registry/table_splitter/table_splitter.go:36-101) is applied HERE, not by tfgpu_plan: the generated name
 * (current table name + splitter + to_string.SerializeToString of the listed columns, "<nil>" for nil / absent values and for items without
 * values) is computed per row from the row image, the rows of a run are grouped by it in order of first appearance and every group goes
 * down on its own (ev.out_table, its own INSERT). It must be the last transformer of the list; `any` / interval columns, []byte outside
 * `string` columns and useLegacyLf are refused.
 * regex_replace_transformer (pkg/transformer/registry/regex_replace/transformer.go:87-142) is applied HERE as well (to the transposed text
 * columns on the host workers, or to the row image before the transposer when items carry column subsets or mixed text types): Go string values of utf8 columns and []byte values of string columns that the step's column filter names are replaced by
 * Regexp.ReplaceAll (see tfgpu_regex_replace_all). No kernel runs regular expressions, so the step (several are fine) must stand at the head
 * of the transformer list; an expression regexp.Compile refuses fails tfgpu_sink_create with TF_E_FATAL_CONFIG like the transformer's
 * constructor, one this library does not carry (\p{..}, (?i) over runes outside ASCII) with TF_E_FATAL_UNSUPPORTED. */
typedef struct tfgpu_sink tfgpu_sink;
#define TF_SINK_EV_ROWS   1   /* one downstream Push of row events of one table */
#define TF_SINK_EV_ITEM   2   /* one downstream Push of a single non-row item */
#define TF_SINK_EV_ERRORS 3   /* errorChangeItems (transformation.go:206-234): the failing input rows, to be pushed with `__transform_error` */
typedef struct tf_sink_event {
    int32_t          type;        /* TF_SINK_EV_* */
    uint32_t         table;       /* index into tf_rows.tables of the push */
    const char*      out_schema;  /* table id after rename_tables */
    const char*      out_table;
    uint64_t         n_items;     /* len(items) of this downstream Push (EV_ROWS: rows that survived the transformers) */
    const uint64_t*  item_idx;    /* the input items behind it: EV_ITEM one entry, EV_ROWS / EV_ERRORS n_items entries */
    const tf_rowerr* errors;      /* EV_ERRORS: n_items entries (row = position in item_idx) */
    const tf_batch*  batch;       /* EV_ROWS with wire_fmt 0 */
    const uint8_t*   wire;        /* EV_ROWS with a wire format: valid until the callback returns */
    uint64_t         wire_len, raw_len, n_frames;
    const uint32_t*  msg_sizes;   /* wire_fmt TF_WIRE_DEBEZIUM: 7 per row — message count, then (key bytes, value bytes | 0xFFFFFFFF tombstone) x 3 */
    int32_t          plan_id;     /* engine plan that produced it, -1 without one */
    int32_t          pad;
} tf_sink_event;
typedef int (*tf_sink_fn)(void* ctx, const tf_sink_event* ev);   /* nonzero return = the downstream Push failed: tfgpu_sink_push returns it */
typedef struct tf_sink_stats {
    uint64_t pushes;                 /* tfgpu_sink_push calls */
    uint64_t downstream_pushes;      /* Push calls that reached the destination */
    uint64_t change_items_pushed;    /* WrapperStats.ChangeItemsPushed */
    uint64_t row_events_pushed;      /* WrapperStats.RowEventsPushed */
    uint64_t inflight_bytes;         /* batchStats' bytes: sum of Size.Values over the counted items */
    uint64_t filter_dropped;         /* MiddlewareFilterStats.Dropped */
    uint64_t transform_dropped;      /* transformation stats Dropped: incoming - transformed */
    uint64_t transform_errors;       /* transformation stats Errors */
    uint64_t max_commit_time;        /* batchStats' `oldestTime` (sic: the reference keeps the LATEST commit time under that name) of the last push */
    uint64_t min_commit_time;        /* its `freshestTime`; both over row + synchronize items with CommitTime != 0 */
    uint64_t without_commit_time;    /* items whose CommitTime is 0 (the reference substitutes time.Now()) */
    uint64_t wire_bytes;             /* bytes handed to the destination */
    uint64_t metering_input_rows;    /* InputDataMetering (pkg/middlewares/metering.go:42-48): items of the tfgpu_sink_push calls that succeeded */
    uint64_t metering_output_rows;   /* OutputDataMetering (metering.go:63-69): items of the downstream pushes that succeeded */
} tf_sink_stats;
int tfgpu_sink_create(tfgpu_engine* e, const char* cfg_json, tfgpu_sink** out);
int tfgpu_sink_destroy(tfgpu_sink* s);
const char* tfgpu_sink_last_error(const tfgpu_sink* s);
int tfgpu_sink_set_callback(tfgpu_sink* s, tf_sink_fn fn, void* ctx);
int tfgpu_sink_set_clickhouse(tfgpu_sink* s, tfgpu_ch_conn* conn);
int tfgpu_sink_push(tfgpu_sink* s, const tf_rows* items);
int tfgpu_sink_stats(const tfgpu_sink* s, tf_sink_stats* out);

/* ------------------------------------------------------------------ N pipelines in one process */
/* Batches dealt round-robin over N sinks (one engine per GPU behind each, created by the caller): every sink's pushes run on a host thread of
 * its own, so the transposes, copies and kernels of consecutive batches overlap across GPUs — there is no collective, rows are independent.
 * What the reference guarantees is the order of a table's rows (transformation.go:131-141 keeps per-table order, not cross-table order): the
 * DELIVERIES of batch k (callback events, ClickHouse INSERTs) start only when every delivery of batch k - 1 has finished, whichever GPU worked
 * on it; the device work of batch k's first run is done by then. tfgpu_dispatcher_submit returns at once (at most two batches queue per sink,
 * like the bufferer's one flush in flight + one collecting, bufferer.go:225-242); `items` and everything it points to must stay valid until
 * tfgpu_dispatcher_wait(seq) returned. The sinks must not be pushed to directly while they belong to a dispatcher. */
typedef struct tfgpu_dispatcher tfgpu_dispatcher;
int tfgpu_dispatcher_create(tfgpu_sink* const* sinks, int n, tfgpu_dispatcher** out);
int tfgpu_dispatcher_submit(tfgpu_dispatcher* d, const tf_rows* items, uint64_t* seq);
int tfgpu_dispatcher_wait(tfgpu_dispatcher* d, uint64_t seq);       /* the return code of that batch's tfgpu_sink_push */
int tfgpu_dispatcher_drain(tfgpu_dispatcher* d);                    /* waits for everything submitted; the first non-zero code, else 0 */
int tfgpu_dispatcher_destroy(tfgpu_dispatcher* d);                  /* drains, stops the threads; the sinks stay the caller's */

/* Host CityHash128 (v1.0.2) as the frames' checksum uses it — exported for the tests' cross-checks against the device and the oracle. */
void tfgpu_host_cityhash128(const uint8_t* p, uint64_t n, uint64_t out[2]);

/* regexp.MustCompile(pattern).ReplaceAll(src, rule) as the regex_replace_transformer step of tfgpu_sink_push computes it
 * (pkg/transformer/registry/regex_replace/transformer.go:127-142; Go's RE2 syntax, leftmost-first, rune-wise): host only, exported for the
 * parity tests and for a shim that wants to validate a transfer's expression up front. Returns the result's length (the bytes are written
 * when they fit `cap`), TF_E_FATAL_CONFIG for an expression regexp.Compile refuses too, TF_E_FATAL_UNSUPPORTED for valid syntax this
 * library does not carry (\p{..} classes, (?i) over runes outside ASCII, oversized programs): such a transfer keeps the Go transformer. */
int64_t tfgpu_regex_replace_all(const char* pattern, const char* rule, const uint8_t* src, uint64_t src_len, uint8_t* dst, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TFGPU_SINK_H_ */
