// Sinker.Push as one call: the reference's sink pipeline below the user's transformers (pkg/sink_factory/sink_factory.go:79-108) over items in
// row form — transformation.Push (pkg/transformer/transformation.go:122-158,236-282) with the device running each table's chain, then
// NonRowSeparator (pkg/middlewares/nonrow_separator.go:29-55), Filter(ExcludeSystemTables) (pkg/middlewares/filter.go:60-77), the Statistician's
// counters (pkg/middlewares/statistician.go:55-68, pkg/stats/sink_wrapper.go:63-78, sink_wrapper_util.go:10-50) and the destination.
// Host-only C++ above the C-ABI of tfgpu.h: it calls tfgpu_plan / tfgpu_push_encode / tfgpu_push_columns like any other client of the library.
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/tfgpu_sink.h"
#include "plan.hpp"

namespace {

struct SinkFail { int rc; std::string msg; };

const char* kind_name(uint8_t k) {            // abstract.Kind strings (kind.go:5-43)
    switch (k) {
    case TF_KIND_INSERT: return "insert"; case TF_KIND_UPDATE: return "update"; case TF_KIND_DELETE: return "delete";
    case TF_KIND_INIT_SHARDED_TABLE_LOAD: return "init_sharded_table_load"; case TF_KIND_INIT_TABLE_LOAD: return "init_load_table";
    case TF_KIND_DONE_TABLE_LOAD: return "done_load_table"; case TF_KIND_DONE_SHARDED_TABLE_LOAD: return "done_sharded_table_load";
    case TF_KIND_DROP_TABLE: return "drop_table"; case TF_KIND_TRUNCATE: return "truncate"; case TF_KIND_DDL: return "DDL";
    case TF_KIND_PG_DDL: return "pg:DDL"; case TF_KIND_SYNCHRONIZE: return ""; default: return "\x01other";
    }
}

// what a transformer does to an item that is not a row event: skip_events may drop it, rename_tables renames it, the rest pass it through
// (filter_rows.go:110, number_to_float.go:59, mask / to_string / to_datetime touch ColumnValues of row events only)
struct HostStep {
    int type = 0;                              // 1 skip_events, 2 rename_tables
    tfplan::NameFilter tables; std::set<std::string> events;
    std::vector<std::pair<std::pair<std::string, std::string>, std::pair<std::string, std::string>>> renames;
};

struct TablePlan {
    int plan_id = -1;
    std::string out_ns, out_name, insert_query;
};

}  // namespace

struct tfgpu_sink {
    tfgpu_engine* e = nullptr;
    std::string err, transformers_json, sink_json, database, debezium_opts;
    bool has_transformers = false, errors_to_sink = true, exclude_sys = true, updateable = false;
    int wire_fmt = 0;
    std::set<std::string> system_tables;
    std::vector<HostStep> host_steps;
    tf_sink_fn fn = nullptr; void* ctx = nullptr; tfgpu_ch_conn* ch = nullptr;
    tfgpu_columnar* pool = nullptr;
    std::map<std::string, TablePlan> plans;
    tf_sink_stats st{};

    TablePlan& plan_for(const tf_table& t) {
        std::string key = std::string(t.schema ? t.schema : "") + '\0' + (t.table ? t.table : "") + '\0' + (t.schema_json ? t.schema_json : "");
        auto it = plans.find(key);
        if (it != plans.end()) return it->second;
        TablePlan tp;
        const std::string ns = t.schema ? t.schema : "", name = t.table ? t.table : "";
        const bool want_sink = wire_fmt == TF_WIRE_CH_NATIVE || wire_fmt == TF_WIRE_CH_NATIVE_LZ4 || wire_fmt == TF_WIRE_CH_JSONEACHROW;
        tfplan::Plan pl;
        try { pl = tfplan::build_plan(ns, name, t.schema_json ? t.schema_json : "[]", transformers_json, want_sink ? sink_json : ""); }
        catch (const tfplan::FatalError& f) { throw SinkFail{f.code, std::string("unable to add table plan: ") + f.what()}; }
        catch (const std::exception& x) { throw SinkFail{TF_E_FATAL_CONFIG, std::string("unable to add table plan: ") + x.what()}; }
        tp.out_ns = pl.out_ns.empty() && pl.out_name.empty() ? ns : pl.out_ns; tp.out_name = pl.out_name.empty() ? name : pl.out_name;
        if (e && (has_transformers || wire_fmt)) {
            const int rc = tfgpu_plan(e, ns.c_str(), name.c_str(), t.schema_json, transformers_json.c_str(), want_sink ? sink_json.c_str() : nullptr, &tp.plan_id);
            if (rc) throw SinkFail{rc, std::string("unable to add table plan: ") + tfgpu_last_error(e)};
        }
        if (ch) {
            std::string cols = "[";
            for (size_t i = 0; i < pl.out_schema.size(); i++) cols += (i ? "," : "") + tfj::quote(pl.out_schema[i].name);
            cols += "]";
            std::vector<char> q(cols.size() + tp.out_name.size() + database.size() + 256);
            const int64_t n = tfgpu_ch_insert_query(database.c_str(), tp.out_name.c_str(), cols.c_str(), updateable, q.data(), q.size());
            if (n < 0) throw SinkFail{(int)n, "cannot build the INSERT statement"};
            tp.insert_query.assign(q.data(), (size_t)n);
        }
        return plans.emplace(std::move(key), std::move(tp)).first->second;
    }

    // a non-row item through the transformer list: false = dropped by skip_events; the table id it leaves with in (ns, name)
    bool host_chain(const std::string& orig_ns, const std::string& orig_name, uint8_t kind, std::string& ns, std::string& name) const {
        ns = orig_ns; name = orig_name;
        for (const HostStep& h : host_steps) {
            if (h.type == 1) { if (tfplan::match_table(h.tables, orig_ns, orig_name) && h.events.count(kind_name(kind))) return false; }
            else {
                bool suitable = false; for (auto& r : h.renames) if (r.first.first == orig_ns && r.first.second == orig_name) suitable = true;   // Suitable on the original id
                if (!suitable) continue;
                for (auto& r : h.renames) if (r.first.first == ns && r.first.second == name) { const auto to = r.second; ns = to.first; name = to.second; }   // last entry wins (map built from the list)
            }
        }
        return true;
    }

    int deliver(const tf_sink_event& ev, const tf_rows* rows, TablePlan* tp) {
        int rc = 0;
        if (ev.type == TF_SINK_EV_ROWS && ch && ev.wire) {
            rc = tfgpu_ch_insert_begin(ch, tp->insert_query.c_str(), "", nullptr);
            if (!rc) rc = tfgpu_ch_insert_data(ch, ev.wire, ev.wire_len);
            if (!rc) rc = tfgpu_ch_insert_end(ch, nullptr, nullptr);
            if (rc) throw SinkFail{rc, std::string("clickhouse: ") + tfgpu_ch_last_error(ch)};
        } else if (fn) {
            rc = fn(ctx, &ev);
            if (rc) throw SinkFail{rc, "the downstream Push failed"};
        }
        // Statistician: counted after the downstream Push succeeded (statistician.go:60-66)
        st.downstream_pushes++; st.change_items_pushed += ev.n_items; st.wire_bytes += ev.wire_len; st.metering_output_rows += ev.n_items;
        if (ev.type != TF_SINK_EV_ITEM) st.row_events_pushed += ev.n_items;
        if (ev.item_idx) for (uint64_t k = 0; k < ev.n_items; k++) {
            const tf_item& it = rows->items[ev.item_idx[k]];
            if (!(TF_KIND_IS_ROW(it.kind) || it.kind == TF_KIND_SYNCHRONIZE)) continue;              // batchStats :16-18
            st.inflight_bytes += it.size_read;
            if (!it.commit_time) { st.without_commit_time++; continue; }
            if (!st.max_commit_time || it.commit_time > st.max_commit_time) st.max_commit_time = it.commit_time;
            if (!st.min_commit_time || it.commit_time < st.min_commit_time) st.min_commit_time = it.commit_time;
        }
        return 0;
    }

    // one maximal run of row events of one (table, schema): transformers + encode on the device, then downstream
    void push_run(const tf_rows* rows, uint32_t table, const std::vector<uint64_t>& idx, TablePlan& tp) {
        const uint64_t n = idx.size();
        if (!n) return;
        tf_sink_event ev{}; ev.table = table; ev.out_schema = tp.out_ns.c_str(); ev.out_table = tp.out_name.c_str(); ev.plan_id = tp.plan_id;
        if (exclude_sys && system_tables.count(tp.out_name)) { st.filter_dropped += n; return; }           // ChangeItem.IsSystemTable looks at Table only
        const tf_batch* batch = nullptr; const tf_row_meta* meta = nullptr; const tf_old_keys* old = nullptr;
        int rc = tfgpu_rows_to_batch(pool, rows, table, idx.data(), n, 0, &batch, &meta, &old);
        if (rc) throw SinkFail{rc, std::string("transpose: ") + tfgpu_columnar_last_error(pool)};
        if (tp.plan_id < 0) {                                                                           // no transformers, columnar hand-over
            ev.type = TF_SINK_EV_ROWS; ev.n_items = n; ev.item_idx = idx.data(); ev.batch = batch;
            deliver(ev, rows, &tp); return;
        }
        tfgpu_result* res = nullptr;
        if (wire_fmt == TF_WIRE_DEBEZIUM)       // queue Debezium serializer: every row kind, OldKeys and the source block's fields from the transposer
            rc = tfgpu_emit_debezium_crud(e, tp.plan_id, debezium_opts.c_str(), batch, old, meta, &res);
        else
            rc = wire_fmt ? tfgpu_push_encode(e, tp.plan_id, wire_fmt, batch, &res) : tfgpu_push_columns(e, tp.plan_id, batch, &res);
        if (rc) throw SinkFail{rc, std::string("device: ") + tfgpu_last_error(e)};
        struct Release { tfgpu_result* r; ~Release() { tfgpu_result_release(r); } } guard{res};
        const uint64_t n_out = tfgpu_result_rows_out(res), n_err = tfgpu_result_n_errors(res);
        st.transform_dropped += n - n_out; st.transform_errors += n_err;
        if (n_err) {
            if (wire_fmt && n_err) {                                                                      // a sink format refuses some rows outright (UPDATE / DELETE kinds): the Go sink must take the run
                const tf_rowerr* er = tfgpu_result_errors(res);
                for (uint64_t k = 0; k < n_err; k++) if (er[k].code == TF_ROWERR_SINK_KIND_HOST || er[k].code == TF_ROWERR_SER_VALUE)
                    throw SinkFail{TF_E_FATAL_UNSUPPORTED, "a row of this run needs the Go sink (update / delete kind or a value the wire format refuses)"};
            }
            if (errors_to_sink) {                                                                         // pushErrors before the transformed items (transformation.go:152-157)
                std::vector<uint64_t> eidx(n_err); std::vector<tf_rowerr> errs(tfgpu_result_errors(res), tfgpu_result_errors(res) + n_err);
                for (uint64_t k = 0; k < n_err; k++) { eidx[k] = idx[errs[k].row]; errs[k].row = (uint32_t)k; }
                tf_sink_event ee = ev; ee.type = TF_SINK_EV_ERRORS; ee.n_items = n_err; ee.item_idx = eidx.data(); ee.errors = errs.data();
                deliver(ee, rows, &tp);
            }
        }
        if (!n_out) return;                                                                              // filter.go:73-75 / an empty Push is not forwarded
        ev.type = TF_SINK_EV_ROWS; ev.n_items = n_out; ev.item_idx = n_out == n ? idx.data() : nullptr;
        if (wire_fmt) { ev.wire = tfgpu_result_bytes(res); ev.wire_len = tfgpu_result_bytes_len(res); ev.raw_len = tfgpu_result_raw_len(res); ev.n_frames = tfgpu_result_n_frames(res); ev.msg_sizes = wire_fmt == TF_WIRE_DEBEZIUM ? tfgpu_result_dbz_msg_sizes(res) : nullptr; }
        else ev.batch = tfgpu_result_batch(res);
        deliver(ev, rows, &tp);
    }

    void push(const tf_rows* rows) {
        st.pushes++; st.max_commit_time = st.min_commit_time = 0;
        // SplitByTableID (utils.go:130-136): groups in order of first appearance, items in input order
        std::vector<std::pair<std::string, std::vector<uint64_t>>> groups; std::map<std::string, size_t> where;
        for (uint64_t i = 0; i < rows->n_items; i++) {
            const tf_item& it = rows->items[i];
            if (it.table >= rows->n_tables) throw SinkFail{TF_E_FATAL_ARG, "item names a table outside tf_rows.tables"};
            const tf_table& t = rows->tables[it.table];
            std::string id = std::string(t.schema ? t.schema : "") + '\0' + (t.table ? t.table : "");
            auto w = where.find(id);
            if (w == where.end()) { w = where.emplace(id, groups.size()).first; groups.emplace_back(id, std::vector<uint64_t>()); }
            groups[w->second].second.push_back(i);
        }
        std::vector<uint64_t> run;
        for (auto& g : groups) {
            uint32_t run_table = 0; TablePlan* run_plan = nullptr;
            auto flush = [&] { if (!run.empty()) { push_run(rows, run_table, run, *run_plan); run.clear(); } };
            for (uint64_t i : g.second) {
                const tf_item& it = rows->items[i];
                const tf_table& t = rows->tables[it.table];
                if (TF_KIND_IS_ROW(it.kind)) {
                    if (!run.empty() && it.table != run_table) flush();                                   // another TableSchema of the same table: its own plan (transformation.go:243-277)
                    if (run.empty()) { run_table = it.table; run_plan = &plan_for(t); }
                    run.push_back(i); continue;
                }
                std::string ns, name;
                if (!host_chain(t.schema ? t.schema : "", t.table ? t.table : "", it.kind, ns, name)) { st.transform_dropped++; continue; }   // the run is NOT cut: the item is gone before NonRowSeparator sees it
                flush();                                                                                  // nonrow_separator.go:38-47
                if (exclude_sys && system_tables.count(name)) { st.filter_dropped++; continue; }
                tf_sink_event ev{}; ev.type = TF_SINK_EV_ITEM; ev.table = it.table; ev.out_schema = ns.c_str(); ev.out_table = name.c_str();
                ev.n_items = 1; ev.item_idx = &i; ev.plan_id = -1;
                deliver(ev, rows, nullptr);
            }
            flush();
        }
    }
};

extern "C" {

int tfgpu_sink_create(tfgpu_engine* e, const char* cfg_json, tfgpu_sink** out) {
    if (!out) return TF_E_FATAL_ARG;
    *out = nullptr;
    auto s = std::make_unique<tfgpu_sink>();
    try {
        tfj::ValuePtr cfg = tfj::parse(cfg_json && *cfg_json ? cfg_json : "{}");
        s->e = e; s->wire_fmt = (int)cfg->get_num("wire_fmt", 0);
        const tfj::Value* trs = cfg->get("transformers");
        s->has_transformers = trs && trs->kind == tfj::Value::Arr && !trs->arr.empty();
        if ((s->has_transformers || s->wire_fmt) && !e) return TF_E_FATAL_NODEVICE;                      // nothing here computes on the CPU
        // the transformer list, re-serialised for tfgpu_plan, and its effect on non-row items
        std::string tj = "[";
        if (s->has_transformers) {
            // the caller's text is passed on verbatim: find the array in cfg_json by re-reading it with the same parser the plans use
            const std::string src = cfg_json; const size_t k = src.find("\"transformers\"");
            size_t a = src.find('[', k), depth = 0, b = a; bool in_str = false;
            for (; b < src.size(); b++) {
                const char c = src[b];
                if (in_str) { if (c == '\\') b++; else if (c == '"') in_str = false; continue; }
                if (c == '"') in_str = true; else if (c == '[') depth++; else if (c == ']' && --depth == 0) break;
            }
            tj = src.substr(a, b - a + 1);
            for (auto& tr : trs->arr) {
                if (tr->kind != tfj::Value::Obj) continue;
                for (auto& kv : tr->obj) {
                    const tfj::Value* c = kv.second.get();
                    if (kv.first == "skip_events") { HostStep h; h.type = 1; h.tables = tfplan::tables_filter(c->get("tables")); for (auto& ev : c->get_str_list("events")) h.events.insert(ev); s->host_steps.push_back(std::move(h)); }
                    else if (kv.first == "rename_tables") {
                        HostStep h; h.type = 2; const tfj::Value* lst = c->get("renameTables");
                        if (lst && lst->kind == tfj::Value::Arr) for (auto& r : lst->arr) {
                            const tfj::Value* o = r->get("originalName"); const tfj::Value* nw = r->get("newName");
                            if (o && nw) h.renames.push_back({{o->get_str("nameSpace"), o->get_str("name")}, {nw->get_str("nameSpace"), nw->get_str("name")}});
                        }
                        s->host_steps.push_back(std::move(h));
                    }
                }
            }
        } else tj = "[]";
        s->transformers_json = tj;
        if (s->wire_fmt == TF_WIRE_DEBEZIUM) {               // the emitter's options are passed on verbatim (tfgpu_emit_debezium's opts_json)
            const std::string src = cfg_json ? cfg_json : ""; const size_t k = src.find("\"debezium\"");
            if (k == std::string::npos) return TF_E_FATAL_CONFIG;
            size_t a = src.find('{', k), depth = 0, b = a; bool in_str = false;
            for (; a != std::string::npos && b < src.size(); b++) {
                const char c = src[b];
                if (in_str) { if (c == '\\') b++; else if (c == '"') in_str = false; continue; }
                if (c == '"') in_str = true; else if (c == '{') depth++; else if (c == '}' && --depth == 0) break;
            }
            if (a == std::string::npos || b >= src.size()) return TF_E_FATAL_CONFIG;
            s->debezium_opts = src.substr(a, b - a + 1);
        }
        s->sink_json = "{\"type\":\"clickhouse\"}";
        s->database = cfg->get_str("database", "default"); s->updateable = cfg->get_bool("updateable", false);
        s->errors_to_sink = cfg->get_str("errors_output", "sink") != "devnull";
        s->exclude_sys = cfg->get_bool("exclude_system_tables", true);
        for (auto& t : cfg->get_str_list("system_tables")) s->system_tables.insert(t);
        const int rc = tfgpu_columnar_create(&s->pool); if (rc) return rc;
    } catch (const std::exception& x) { return TF_E_FATAL_CONFIG; }
    *out = s.release();
    return TF_OK;
}

int tfgpu_sink_destroy(tfgpu_sink* s) { if (!s) return TF_E_FATAL_ARG; if (s->pool) tfgpu_columnar_destroy(s->pool); delete s; return TF_OK; }
const char* tfgpu_sink_last_error(const tfgpu_sink* s) { return s ? s->err.c_str() : "null sink"; }
int tfgpu_sink_set_callback(tfgpu_sink* s, tf_sink_fn fn, void* ctx) { if (!s) return TF_E_FATAL_ARG; s->fn = fn; s->ctx = ctx; return TF_OK; }
int tfgpu_sink_set_clickhouse(tfgpu_sink* s, tfgpu_ch_conn* conn) {
    if (!s) return TF_E_FATAL_ARG;
    if (conn && s->wire_fmt != TF_WIRE_CH_NATIVE_LZ4 && s->wire_fmt != TF_WIRE_CH_NATIVE) { s->err = "the ClickHouse writer takes wire_fmt TF_WIRE_CH_NATIVE_LZ4 (compression on) or TF_WIRE_CH_NATIVE"; return TF_E_FATAL_CONFIG; }
    s->ch = conn; s->plans.clear(); return TF_OK;
}

int tfgpu_sink_push(tfgpu_sink* s, const tf_rows* items) {
    if (!s || !items) return TF_E_FATAL_ARG;
    try { s->push(items); s->st.metering_input_rows += items->n_items; return TF_OK; }
    catch (const SinkFail& f) { s->err = f.msg; return f.rc; }
    catch (const std::bad_alloc&) { s->err = "host allocation failed"; return TF_E_RETRY_OOM; }
    catch (const std::exception& x) { s->err = x.what(); return TF_E_FATAL_CONFIG; }
}

int tfgpu_sink_stats(const tfgpu_sink* s, tf_sink_stats* out) { if (!s || !out) return TF_E_FATAL_ARG; *out = s->st; return TF_OK; }

}  // extern "C"
