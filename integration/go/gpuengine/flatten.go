// Package gpuengine is the cgo shim a maintainer adds to transferia to bind libtfgpu.so (see INTEGRATION.md).
// It cannot be compiled in the image this repository was built in (no Go toolchain): it documents, completely, how the
// C-ABI of include/tfgpu.h and include/tfgpu_sink.h is driven from the reference's side.
package gpuengine

/*
#cgo LDFLAGS: -ltfgpu
#include "tfgpu_sink.h"
*/
import "C"

import (
	"encoding/binary"
	"encoding/json"
	"math"
	"time"
	"unsafe"

	"github.com/transferia/transferia/pkg/abstract"
)

// flattener turns []abstract.ChangeItem into the row form of include/tfgpu_sink.h: one tf_item per ChangeItem and one byte image of
// the boxed values (a tag byte naming the Go dynamic type, then the payload) — plain appends, one type switch per value, no cgo call
// inside the loop. The buffers are reused across Push calls.
type flattener struct {
	items   []C.tf_item
	vals    []byte
	strs    []byte
	tables  []C.tf_table
	tableOf map[tableKey]uint32
	keep    []unsafe.Pointer // C strings of the table descriptors, freed on Close
}

type tableKey struct {
	id   abstract.TableID
	hash string // TableSchema.Hash(): items of one table may change schema inside a batch (transformation.go:243-277)
}

func (f *flattener) reset() { f.items, f.vals, f.strs = f.items[:0], f.vals[:0], f.strs[:0] }

func (f *flattener) table(it *abstract.ChangeItem) uint32 {
	h, _ := it.TableSchema.Hash()
	k := tableKey{it.TableID(), h}
	if idx, ok := f.tableOf[k]; ok {
		return idx
	}
	sj, _ := json.Marshal(it.TableSchema.Columns()) // the ColSchema JSON tags of col_schema.go:14-29
	t := C.tf_table{schema: C.CString(it.Schema), table: C.CString(it.Table), schema_json: C.CString(string(sj))}
	f.keep = append(f.keep, unsafe.Pointer(t.schema), unsafe.Pointer(t.table), unsafe.Pointer(t.schema_json))
	f.tables = append(f.tables, t)
	f.tableOf[k] = uint32(len(f.tables) - 1)
	return uint32(len(f.tables) - 1)
}

var kindCode = map[abstract.Kind]C.uint8_t{
	abstract.InsertKind: C.TF_KIND_INSERT, abstract.UpdateKind: C.TF_KIND_UPDATE, abstract.DeleteKind: C.TF_KIND_DELETE,
	abstract.InitShardedTableLoad: C.TF_KIND_INIT_SHARDED_TABLE_LOAD, abstract.InitTableLoad: C.TF_KIND_INIT_TABLE_LOAD,
	abstract.DoneTableLoad: C.TF_KIND_DONE_TABLE_LOAD, abstract.DoneShardedTableLoad: C.TF_KIND_DONE_SHARDED_TABLE_LOAD,
	abstract.DropTableKind: C.TF_KIND_DROP_TABLE, abstract.TruncateTableKind: C.TF_KIND_TRUNCATE, abstract.DDLKind: C.TF_KIND_DDL,
	abstract.PgDDLKind: C.TF_KIND_PG_DDL, abstract.SynchronizeKind: C.TF_KIND_SYNCHRONIZE,
}

// value appends one boxed value. The tags are the TF_V_* constants of tfgpu_sink.h.
func (f *flattener) value(v any) {
	le := binary.LittleEndian
	switch x := v.(type) {
	case nil:
		f.vals = append(f.vals, C.TF_V_NIL)
	case bool:
		b := byte(0)
		if x {
			b = 1
		}
		f.vals = append(f.vals, C.TF_V_BOOL, b)
	case int8:
		f.vals = append(f.vals, C.TF_V_INT8, byte(x))
	case int16:
		f.vals = le.AppendUint16(append(f.vals, C.TF_V_INT16), uint16(x))
	case int32:
		f.vals = le.AppendUint32(append(f.vals, C.TF_V_INT32), uint32(x))
	case int64:
		f.vals = le.AppendUint64(append(f.vals, C.TF_V_INT64), uint64(x))
	case int:
		f.vals = le.AppendUint64(append(f.vals, C.TF_V_INT64), uint64(int64(x)))
	case uint8:
		f.vals = append(f.vals, C.TF_V_UINT8, x)
	case uint16:
		f.vals = le.AppendUint16(append(f.vals, C.TF_V_UINT16), x)
	case uint32:
		f.vals = le.AppendUint32(append(f.vals, C.TF_V_UINT32), x)
	case uint64:
		f.vals = le.AppendUint64(append(f.vals, C.TF_V_UINT64), x)
	case float32:
		f.vals = le.AppendUint32(append(f.vals, C.TF_V_FLOAT32), math.Float32bits(x))
	case float64:
		f.vals = le.AppendUint64(append(f.vals, C.TF_V_FLOAT64), math.Float64bits(x))
	case string:
		f.vals = append(le.AppendUint32(append(f.vals, C.TF_V_STRING), uint32(len(x))), x...)
	case []byte:
		f.vals = append(le.AppendUint32(append(f.vals, C.TF_V_BYTES), uint32(len(x))), x...)
	case time.Time:
		f.vals = le.AppendUint32(le.AppendUint64(append(f.vals, C.TF_V_TIME), uint64(x.Unix())), uint32(x.Nanosecond()))
	case time.Duration:
		f.vals = le.AppendUint64(append(f.vals, C.TF_V_DURATION), uint64(x))
	case json.Number:
		f.vals = append(le.AppendUint32(append(f.vals, C.TF_V_JSONNUM), uint32(len(x))), x...)
	default: // maps, slices, ...: their json.Marshal text (what columntypes.Restore / the serializers would marshal)
		b, _ := json.Marshal(x)
		f.vals = append(le.AppendUint32(append(f.vals, C.TF_V_JSON), uint32(len(b))), b...)
	}
}

func (f *flattener) add(it *abstract.ChangeItem) {
	var ci C.tf_item
	ci.lsn, ci.commit_time, ci.size_read, ci.size_values = C.uint64_t(it.LSN), C.uint64_t(it.CommitTime), C.uint64_t(it.Size.Read), C.uint64_t(it.Size.Values)
	ci.id, ci.counter, ci.table = C.uint32_t(it.ID), C.int32_t(it.Counter), C.uint32_t(f.table(it))
	if k, ok := kindCode[it.Kind]; ok {
		ci.kind = k
	} else {
		ci.kind = C.TF_KIND_OTHER
	}
	ci.txid_off, ci.txid_len = C.uint32_t(len(f.strs)), C.uint32_t(len(it.TxID))
	f.strs = append(f.strs, it.TxID...)
	ci.part_off, ci.part_len = C.uint32_t(len(f.strs)), C.uint32_t(len(it.PartID))
	f.strs = append(f.strs, it.PartID...)
	ci.values_off, ci.n_values = C.uint64_t(len(f.vals)), C.uint32_t(len(it.ColumnValues))
	cols := it.TableSchema.Columns()
	dense := len(it.ColumnNames) == len(cols)
	for i := 0; dense && i < len(cols); i++ {
		dense = it.ColumnNames[i] == cols[i].ColumnName
	}
	if !dense { // a toasted update / a column subset: every value carries its schema index
		ci.flags = C.TF_ITEM_SPARSE
		pos := abstract.MakeMapColNameToIndex(cols)
		for i, name := range it.ColumnNames {
			f.vals = binary.LittleEndian.AppendUint16(f.vals, uint16(pos[name]))
			f.value(it.ColumnValues[i])
		}
	} else {
		for _, v := range it.ColumnValues {
			f.value(v)
		}
	}
	ci.old_keys_off = C.uint64_t(math.MaxUint64)
	if len(it.OldKeys.KeyNames) > 0 { // old_keys.go:3-7
		ci.old_keys_off = C.uint64_t(len(f.vals))
		f.vals = binary.LittleEndian.AppendUint16(f.vals, uint16(len(it.OldKeys.KeyNames)))
		pos := abstract.MakeMapColNameToIndex(cols)
		for i, name := range it.OldKeys.KeyNames {
			f.vals = binary.LittleEndian.AppendUint16(f.vals, uint16(pos[name]))
			f.value(it.OldKeys.KeyValues[i])
		}
	}
	f.items = append(f.items, ci)
}

// rows points a tf_rows at the flattener's buffers; the caller pins them for the duration of the C call.
func (f *flattener) rows() C.tf_rows {
	var r C.tf_rows
	r.n_items, r.n_tables = C.uint64_t(len(f.items)), C.uint32_t(len(f.tables))
	if len(f.items) > 0 {
		r.items = &f.items[0]
	}
	if len(f.tables) > 0 {
		r.tables = &f.tables[0]
	}
	if len(f.vals) > 0 {
		r.values = (*C.uint8_t)(unsafe.Pointer(&f.vals[0]))
	}
	r.values_len = C.uint64_t(len(f.vals))
	if len(f.strs) > 0 {
		r.strings = (*C.uint8_t)(unsafe.Pointer(&f.strs[0]))
	}
	r.strings_len = C.uint64_t(len(f.strs))
	return r
}
