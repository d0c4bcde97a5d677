package gpuengine

/*
#include "tfgpu_sink.h"
extern int goSinkEvent(void* ctx, tf_sink_event* ev);   // //export below
static int sink_trampoline(void* ctx, const tf_sink_event* ev) { return goSinkEvent(ctx, (tf_sink_event*)ev); }
static void sink_install(tfgpu_sink* s, void* ctx) { tfgpu_sink_set_callback(s, sink_trampoline, ctx); }
*/
import "C"

import (
	"encoding/json"
	"net"
	"runtime"
	"runtime/cgo"
	"unsafe"

	"github.com/transferia/transferia/library/go/core/xerrors"
	"github.com/transferia/transferia/pkg/abstract"
	"github.com/transferia/transferia/pkg/abstract/model"
	"github.com/transferia/transferia/pkg/middlewares"
)

// Sinker is the abstract.Sinker the pluggable-transformer hook returns (pkg/middlewares/pluggable_transformer.go:19-30): it owns one engine
// (one GPU), one tfgpu_sink (the reference's pipeline below the user transformers as one C call) and, for a ClickHouse destination, one
// native-protocol connection the device's LZ4 frames leave through.
type Sinker struct {
	eng   *C.tfgpu_engine
	sink  *C.tfgpu_sink
	ch    *C.tfgpu_ch_conn
	conn  *net.TCPConn
	next  abstract.Sinker // the destination Sinker for what the native writer does not take: control items, error rows
	flat  flattener
	batch []abstract.ChangeItem // the items of the Push in flight (the callback indexes into it)
	self  cgo.Handle
}

type Config struct {
	Device        int
	Transformers  json.RawMessage // transfer YAML `transformation.transformers`, as JSON
	ErrorsOutput  string          // "sink" | "devnull" (transformation.go:183-205)
	SystemTables  []string        // abstract.SystemTables()
	Database      string
	WireFmt       int    // C.TF_WIRE_CH_NATIVE_LZ4, 0 (columnar hand-over), C.TF_WIRE_DEBEZIUM ...
	ClickHouse    string // host:port of the native endpoint, "" = no native writer
	User, Pass    string
}

func classify(rc C.int, msg string) error {
	if rc == 0 {
		return nil
	}
	err := xerrors.Errorf("tfgpu rc=%d: %s", int(rc), msg)
	if rc < 0 {
		return abstract.NewFatalError(err) // pkg/abstract/errors.go:14; rc > 0 stays retriable (middlewares/retrier.go:46-80)
	}
	return err
}

func New(cfg Config, next abstract.Sinker) (*Sinker, error) {
	s := &Sinker{next: next, flat: flattener{tableOf: map[tableKey]uint32{}}}
	dev := C.int(cfg.Device)
	if rc := C.tfgpu_engine_create(nil, &dev, 1, &s.eng); rc != 0 {
		return nil, classify(rc, "tfgpu_engine_create (no CPU fallback)")
	}
	j, _ := json.Marshal(map[string]any{"transformers": cfg.Transformers, "errors_output": cfg.ErrorsOutput, "system_tables": cfg.SystemTables,
		"exclude_system_tables": true, "database": cfg.Database, "wire_fmt": cfg.WireFmt})
	cj := C.CString(string(j))
	defer C.free(unsafe.Pointer(cj))
	if rc := C.tfgpu_sink_create(s.eng, cj, &s.sink); rc != 0 {
		return nil, classify(rc, "tfgpu_sink_create")
	}
	s.self = cgo.NewHandle(s)
	C.sink_install(s.sink, unsafe.Pointer(&s.self))
	if cfg.ClickHouse != "" {
		c, err := net.Dial("tcp", cfg.ClickHouse)
		if err != nil {
			return nil, xerrors.Errorf("dial clickhouse: %w", err)
		}
		s.conn = c.(*net.TCPConn)
		f, _ := s.conn.File() // a duplicate descriptor in blocking mode: the C side reads and writes it with poll timeouts
		opts, _ := json.Marshal(map[string]any{"database": cfg.Database, "user": cfg.User, "password": cfg.Pass, "compression": true})
		co := C.CString(string(opts))
		defer C.free(unsafe.Pointer(co))
		if rc := C.tfgpu_ch_open(C.int(f.Fd()), co, &s.ch); rc != 0 {
			return nil, classify(rc, C.GoString(C.tfgpu_ch_last_error(s.ch)))
		}
		C.tfgpu_sink_set_clickhouse(s.sink, s.ch)
	}
	return s, nil
}

// Push: Sinker.Push([]abstract.ChangeItem) (pkg/abstract/sink.go:14-19). Everything below the flatten happens inside tfgpu_sink_push.
func (s *Sinker) Push(items []abstract.ChangeItem) error {
	s.flat.reset()
	for i := range items {
		s.flat.add(&items[i])
	}
	rows := s.flat.rows()
	var pin runtime.Pinner // the Go-allocated buffers tf_rows points into stay put during the call
	defer pin.Unpin()
	for _, p := range []any{&s.flat.items, &s.flat.vals, &s.flat.strs, &s.flat.tables} {
		pin.Pin(p)
	}
	s.batch = items
	rc := C.tfgpu_sink_push(s.sink, &rows)
	s.batch = nil
	return classify(rc, C.GoString(C.tfgpu_sink_last_error(s.sink)))
}

//export goSinkEvent
func goSinkEvent(ctx unsafe.Pointer, ev *C.tf_sink_event) C.int {
	s := (*(*cgo.Handle)(ctx)).Value().(*Sinker)
	idx := unsafe.Slice((*uint64)(unsafe.Pointer(ev.item_idx)), int(ev.n_items))
	switch ev._type {
	case C.TF_SINK_EV_ITEM: // control items travel alone (NonRowSeparator): hand the original item on, renamed if the chain renamed it
		it := s.batch[idx[0]]
		it.Schema, it.Table = C.GoString(ev.out_schema), C.GoString(ev.out_table)
		if err := s.next.Push([]abstract.ChangeItem{it}); err != nil {
			return 1
		}
	case C.TF_SINK_EV_ERRORS: // transformation.pushErrors -> errorChangeItems (transformation.go:206-234)
		errs := unsafe.Slice((*C.tf_rowerr)(unsafe.Pointer(ev.errors)), int(ev.n_items))
		out := make([]abstract.ChangeItem, 0, len(errs))
		for k, e := range errs {
			out = append(out, errorItem(s.batch[idx[k]], rowErrorText(uint16(e.code)))) // adds the `__transform_error` column
		}
		if err := s.next.Push(out); err != nil {
			return 1
		}
	case C.TF_SINK_EV_ROWS: // only without a native writer: encoded bytes (ev.wire) or columnar rows (ev.batch -> tfgpu_batch_to_rows)
		if err := s.forwardRows(ev); err != nil {
			return 1
		}
	}
	return 0
}

func (s *Sinker) Close() error {
	if s.ch != nil {
		C.tfgpu_ch_close(s.ch)
	}
	if s.conn != nil {
		_ = s.conn.Close()
	}
	C.tfgpu_sink_destroy(s.sink)
	C.tfgpu_engine_destroy(s.eng)
	s.self.Delete()
	return s.next.Close()
}

// registration: same pattern as registry/batch_splitter/plugable_transformer.go:14-75
func init() {
	middlewares.PlugTransformer(func(t *model.Transfer, _ any, _ any) func(abstract.Sinker) abstract.Sinker {
		cfg, ok := configFor(t) // runtime flag + destination type; nil when the transfer is not eligible
		if !ok {
			return nil
		}
		return func(next abstract.Sinker) abstract.Sinker {
			s, err := New(cfg, next)
			if err != nil {
				return next // engine unavailable: the stock CPU pipeline stays in place
			}
			return s
		}
	})
}
