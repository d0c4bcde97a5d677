"""ClickHouse native client writer (include/tfgpu_sink.h) against an independent server-side peer over a socketpair:
Hello / addendum / Query / Data packets are decoded field by field by tests/ch_peer.py, the frame streams that arrive are decoded with the
oracle (and with pyarrow's LZ4 + the independent CityHash128) and must reconstruct the oracle's native block byte for byte."""
import os
import socket
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ch_peer
from cityhash_independent import cityhash128
from transferia_b200 import abi, engine, sink, workload


def _pair():
    a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    return a, b


def _headline(po, nrows):
    batch, schema = workload.make_hits_batch(nrows)
    k = workload.counterid_threshold(batch, schema)
    trs = workload.headline_transformers(k)
    plan = po.build_plan("public", "hits", schema, trs)
    return batch, schema, trs, plan


def _result_columns(plan):
    return [(c["name"], c["ch_type"]) for c in plan.describe()["result_schema"]] if hasattr(plan, "describe") else None


def test_host_cityhash_equals_independent_and_oracle(po):
    rng = np.random.default_rng(7)
    for n in list(range(0, 40)) + [63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 4096, 30729]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert sink.host_cityhash128(data) == cityhash128(data), n
        assert sink.host_cityhash128(data) == po.cityhash128(data), n


def test_insert_query_text():
    # sink_table.go:633-660: backquoted names, `?` placeholders cut at VALUES by the driver; updateable sinks add the two system columns
    assert sink.insert_query("db", "hits", ["a", "b c"]) == "INSERT INTO `db`.`hits` (`a`,`b c`) VALUES"
    assert sink.insert_query("db", "t", ["x"], updateable=True) == \
        "INSERT INTO `db`.`t` (`x`,`__data_transfer_commit_time`,`__data_transfer_delete_time`) VALUES"
    # the statement sqlmock expects in sink_table_test.go:126-127 (TestTable_doOperation_updatable_delete), up to VALUES where the driver cuts it
    cols = ["id", "old_status", "new_status", "claim_id", "event_time", "comment", "ticket", "old_current_point", "new_current_point"]
    assert sink.insert_query("db", "test_table", cols, updateable=True) == \
        "INSERT INTO `db`.`test_table` (`id`,`old_status`,`new_status`,`claim_id`,`event_time`,`comment`,`ticket`,`old_current_point`,`new_current_point`," \
        "`__data_transfer_commit_time`,`__data_transfer_delete_time`) VALUES"


def test_insert_exchange_with_oracle_frames(po):
    batch, schema, trs, plan = _headline(po, 6000)
    ref = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE_LZ4, 30720)
    ref2 = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE_LZ4, 4096)
    cols = [(f"c{i}", "Int32") for i in range(3)]
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, cols, expect_raw_len=[len(ref.raw), len(ref2.raw)])
    peer.start()
    w = sink.ClickHouseWriter(cli, database="analytics", user="loader", password="s3cret", read_timeout_ms=20000)
    info = w.server_info
    assert info["revision"] == 54460 and info["server_revision"] == ch_peer.SERVER_REVISION and info["timezone"] == "Europe/Amsterdam"
    assert info["display_name"] == "ch-test-1" and info["patch"] == 7
    q = sink.insert_query("analytics", "hits", [n for n, _ in cols])
    sample = w.prepare_batch(q, query_id="chv2-streamer-1", settings={"insert_distributed_sync": "1", "max_insert_threads": 4})
    assert sample == [{"name": n, "type": t} for n, t in cols]
    w.append_frames(ref.wire)
    w.append_frames(ref2.wire)
    wrote = w.send()
    peer.join(20)
    assert peer.error is None, peer.error
    assert wrote == (1234, 12340)
    assert peer.hello == {"name": "transferia-tfgpu", "major": 2, "minor": 46, "revision": 54460, "database": "analytics", "user": "loader",
                          "password": "s3cret", "quota_key": ""}
    qq = peer.queries[0]
    assert qq["query_id"] == "chv2-streamer-1" and qq["body"] == q and qq["stage"] == 2 and qq["compression"] == 1 and qq["kind"] == 1
    assert qq["interface"] == 1 and qq["client_revision"] == 54460 and qq["initial_user"] == "loader" and qq["parameters"] == []
    assert qq["settings"] == {"insert_distributed_sync": (0, "1"), "max_insert_threads": (0, "4")}
    got = peer.blocks[0]
    assert got == [ref.wire, ref2.wire]                       # forwarded untouched, packet headers aside
    for wire, want in zip(got, (ref.raw, ref2.raw)):
        raw, _ = po.ch_decode_frames(wire)
        assert raw == want
    # first frame also through the independent checksum + pyarrow LZ4
    first = got[0][:16 + struct.unpack("<I", got[0][17:21])[0]]
    assert ch_peer.frame_payload(first) == ref.raw[:len(ch_peer.frame_payload(first))]
    st = w.stats()
    assert st["data_packets"] == 2 and st["bytes_out"] > len(ref.wire) + len(ref2.wire)
    w.close(); cli.close()


def test_two_inserts_on_one_connection_and_quiet_server(po):
    batch, schema, trs, plan = _headline(po, 1500)
    ref = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE_LZ4, 8192)
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, [("a", "String")], expect_raw_len=[len(ref.raw)], n_inserts=2, chatter=False)
    peer.start()
    w = sink.ClickHouseWriter(cli, read_timeout_ms=20000)
    for k in range(2):                                         # streamer.restart(): Send, then PrepareBatch again on the same connection
        w.prepare_batch("INSERT INTO `d`.`t` (`a`) VALUES", query_id=f"q{k}")
        w.append_frames(ref.wire)
        assert w.send() == (1234 + k, (1234 + k) * 10)
    peer.join(20)
    assert peer.error is None, peer.error
    assert [b[0] for b in peer.blocks] == [ref.wire, ref.wire]
    w.close(); cli.close()


def test_uncompressed_connection_sends_the_raw_block(po):
    """compression off (clickhouse_go.Options.Compression nil): the Query packet says so, the Data packet body is the native block itself
    (TF_WIRE_CH_NATIVE) and the server's sample block comes back uncompressed too."""
    batch, schema, trs, plan = _headline(po, 800)
    ref = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE)
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, [("a", "Int64"), ("b", "Nullable(String)")], expect_raw_len=[len(ref.raw)]); peer.start()
    w = sink.ClickHouseWriter(cli, compression=False, read_timeout_ms=20000)
    assert w.prepare_batch("INSERT INTO `d`.`t` (`a`,`b`) VALUES") == [{"name": "a", "type": "Int64"}, {"name": "b", "type": "Nullable(String)"}]
    w.append_frames(ref.raw)
    assert w.send() == (1234, 12340)
    peer.join(20)
    assert peer.error is None, peer.error
    assert peer.queries[0]["compression"] == 0 and peer.blocks[0] == [ref.raw]
    w.close(); cli.close()


def test_server_exception_and_auth_failure():
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, [("a", "String")], fail_hello=True); peer.start()
    with pytest.raises(engine.EngineError) as ei:
        sink.ClickHouseWriter(cli, read_timeout_ms=5000)
    assert ei.value.rc == 4 and ei.value.retriable and "Authentication failed" in str(ei.value) and ei.value.exception_code == 516
    peer.join(5); cli.close()
    # exception instead of the sample block (unknown table), then after the data (e.g. a constraint / memory limit)
    for when, code in ((-60, 60), (241, 241)):
        cli, srv = _pair()
        peer = ch_peer.Peer(srv, [("a", "String")], fail_insert_with=(when, "boom"), expect_raw_len=[]); peer.start()
        w = sink.ClickHouseWriter(cli, read_timeout_ms=5000)
        with pytest.raises(engine.EngineError) as ei:
            w.prepare_batch("INSERT INTO `d`.`nope` (`a`) VALUES")
            w.send()
        assert ei.value.rc == 4 and ei.value.exception_code == code and "boom" in str(ei.value)
        peer.join(5); w.close(); cli.close()


def test_closed_peer_is_a_retriable_io_error():
    cli, srv = _pair()
    srv.close()
    with pytest.raises(engine.EngineError) as ei:
        sink.ClickHouseWriter(cli, read_timeout_ms=2000)
    assert ei.value.rc == 3 and ei.value.retriable
    cli.close()


def test_old_server_is_refused():
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, [("a", "String")], revision=54451); peer.start()
    with pytest.raises(engine.EngineError) as ei:
        sink.ClickHouseWriter(cli, read_timeout_ms=2000)
    assert ei.value.rc == -2 and "54451" in str(ei.value)
    cli.close(); peer.join(5)


@pytest.mark.gpu
def test_device_frames_reach_the_peer_full_size(eng, po):
    """1 M x 99 rows: the device's frame stream goes out through the writer straight from the engine's pinned landing buffer; the peer's
    copy decodes (oracle decoder: checksums + LZ4) to the oracle's block."""
    batch, schema, trs, plan = _headline(po, 1_000_000)
    pid = eng.plan("public", "hits", schema, trs, {"type": "clickhouse"})
    res = eng.push_encode(pid, batch, abi.TF_WIRE_CH_NATIVE_LZ4)
    ref = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE, eng.frame_bytes)
    names = [c["name"] for c in eng.describe(pid)["result_schema"]]
    cli, srv = _pair()
    peer = ch_peer.Peer(srv, [(n, "String") for n in names], expect_raw_len=[len(ref.raw)], chatter=True); peer.start()
    w = sink.ClickHouseWriter(cli, read_timeout_ms=120000)
    w.prepare_batch(sink.insert_query("default", "hits", names))
    w.append_frames(res.wire)
    w.send()
    peer.join(120)
    assert peer.error is None, peer.error
    raw, nframes = po.ch_decode_frames(peer.blocks[0][0])
    assert raw == ref.raw and nframes == res.n_frames
    w.close(); cli.close()


def test_client_survives_garbage_from_the_server():
    """Whatever bytes a broken or hostile peer sends (random data, truncated packets, absurd lengths, a frame whose checksum or LZ4 body is wrong),
    the client answers with an error code — never a crash, never a hang past its read timeout."""
    rng = np.random.default_rng(99)
    hello_ok = (ch_peer.uvarint(0) + ch_peer.string("ClickHouse") + ch_peer.uvarint(24) + ch_peer.uvarint(3) + ch_peer.uvarint(54467) +
                ch_peer.string("UTC") + ch_peer.string("x") + ch_peer.uvarint(1))
    sample = ch_peer.block([("a", "String", b"")], 0)
    good_frame = ch_peer.compress_frame(sample)
    bad_sum = bytes([good_frame[0] ^ 1]) + good_frame[1:]
    bad_lz4 = bytearray(ch_peer.compress_frame(b"A" * 4000)); bad_lz4[30] ^= 0xff
    # recompute nothing: the body no longer matches its checksum either way; a frame with a valid checksum over a broken body:
    import struct as st
    body = b"\xf0" + b"\xff" * 40                                             # literal length runs past the end
    f = b"\x82" + st.pack("<II", len(body) + 9, 100) + body; lo, hi = cityhash128(f); broken_body = st.pack("<QQ", lo, hi) + f
    answers = [b"", b"\x00", hello_ok[:7], rng.integers(0, 256, 64, dtype=np.uint8).tobytes(), b"\xff" * 40,
               hello_ok + rng.integers(0, 256, 200, dtype=np.uint8).tobytes(),
               hello_ok + ch_peer.uvarint(1) + ch_peer.string("") + bad_sum,
               hello_ok + ch_peer.uvarint(1) + ch_peer.string("") + bytes(bad_lz4),
               hello_ok + ch_peer.uvarint(1) + ch_peer.string("") + broken_body,
               hello_ok + ch_peer.uvarint(1) + ch_peer.uvarint(2 ** 40),                             # a string of a terabyte
               hello_ok + ch_peer.uvarint(10) + ch_peer.string("") + ch_peer.block([("m", "Map(String, String)", b"")], 3),   # a Log block with a type the reader cannot skip
               hello_ok + ch_peer.uvarint(99)]
    for _ in range(40):
        n = int(rng.integers(1, 120)); answers.append(hello_ok + rng.integers(0, 20, n, dtype=np.uint8).tobytes())
    import threading
    for ans in answers:
        cli, srv = _pair()
        def serve(s=srv, a=ans):
            try:
                s.recv(4096)
                if a: s.sendall(a)
                s.settimeout(0.3)
                try:
                    while s.recv(65536): pass
                except OSError: pass
            finally:
                s.close()
        th = threading.Thread(target=serve, daemon=True); th.start()
        try:
            w = sink.ClickHouseWriter(cli, read_timeout_ms=400)
            try:
                w.prepare_batch("INSERT INTO `d`.`t` (`a`) VALUES")
                w.append_frames(good_frame); w.send()
                raise AssertionError("garbage was accepted as a complete INSERT exchange")
            except engine.EngineError as ex:
                assert ex.rc in (3, 4, -2, -5), ex.rc
            w.close()
        except engine.EngineError as ex:
            assert ex.rc in (3, 4, -2, -5), ex.rc
        cli.close(); th.join(3)
