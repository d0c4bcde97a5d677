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
 * The fd stays owned by the caller; tfgpu_ch_close only frees the handle. */
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

/* Host CityHash128 (v1.0.2) as the frames' checksum uses it — exported for the tests' cross-checks against the device and the oracle. */
void tfgpu_host_cityhash128(const uint8_t* p, uint64_t n, uint64_t out[2]);

#ifdef __cplusplus
}
#endif
#endif /* TFGPU_SINK_H_ */
