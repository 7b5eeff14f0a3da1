"""The DEVICE-PARSE path on the CPU: broadway_b200/csrc/kp_core.h — the body of CUDA kernel Kp, the device-side
slice_data()/CAVLC parser — compiled as plain C++ with one lane per warp (oracle/kp_cpu.cpp) under the product's own
host decoder running in device-parse mode (NAL scan, slice headers, DPB on the host; slices queued as picture blocks,
include/h264b200_slices.h; pictures ended by the next access unit).  Checked here without a GPU:

  * every golden case and every packet-loss case decodes to the reference's frames (and concealed-macroblock counts);
  * the records and coefficient slots Kp's code produces are byte-identical to the host parser's, picture by picture;
  * bit-flipped streams decode to the same pictures whichever side parses (same error detection, same concealment);
  * the multi-stream runner's device-parse pipeline (look-ahead scan, output FIFO, release) delivers golden frames.

The GPU tests (tests/test_parity_gpu.py) repeat the first three through the real kernel.
"""
import ctypes
import hashlib
import json
import os
import random

import pytest

import cases
import util
from broadway_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "loss.json")))
ALL = cases.SMALL + cases.FULL[:1]


def _device_md5(data, monkeypatch):
    monkeypatch.setenv("H264B200_PARSE", "device")
    return util.oracle_md5(data)


@pytest.mark.parametrize("case", ALL, ids=[c[0] for c in ALL])
def test_device_parse_matches_reference_golden(case, golden, monkeypatch):
    md5s, summary = _device_md5(cases.make_stream(case), monkeypatch)
    assert summary["err_mbs"] == 0
    assert md5s == golden[case[0]]["frame_md5"]


@pytest.mark.parametrize("lc", cases.LOSS, ids=[c[0] for c in cases.LOSS])
def test_device_parse_concealment_matches_reference_golden(lc, monkeypatch):
    g = LOSS_GOLDEN[lc[0]]
    md5s, summary = _device_md5(cases.make_loss_stream(lc), monkeypatch)
    assert md5s == g["frame_md5"]
    assert summary["err_mbs"] == g["err_mbs"]


def test_device_parse_arbitrary_slice_order(golden, monkeypatch):
    for name in ("multi_slice", "deblock_idc2", "odd_size"):
        case = next(c for c in cases.SMALL if c[0] == name)
        md5s, summary = _device_md5(cases.reverse_slice_order(cases.make_stream(case)), monkeypatch)
        assert summary["err_mbs"] == 0 and md5s == golden[name]["frame_md5"]


def test_bit_errors_device_parse_equals_host_parse(monkeypatch):
    """Both parsers are the same algorithm: whatever a corrupted stream decodes to, it is the same pictures."""
    rng = random.Random(202)
    bases = [cases.make_stream(c) for c in cases.SMALL[:14]]
    for _ in range(40):
        data = bytearray(rng.choice(bases))
        for _ in range(rng.randrange(1, 4)):
            data[rng.randrange(60, len(data))] ^= 1 << rng.randrange(8)
        monkeypatch.setenv("H264B200_PARSE", "host")
        a, sa = util.oracle_md5(bytes(data))
        monkeypatch.setenv("H264B200_PARSE", "device")
        b, sb = util.oracle_md5(bytes(data))
        assert a == b and sa["err_mbs"] == sb["err_mbs"]


# --------------------------------------------------------------------------- records, byte for byte
@pytest.fixture(scope="module")
def L():
    return util.cpuchk_lib()


capture_records = util.capture_records


@pytest.mark.parametrize("case", ALL, ids=[c[0] for c in ALL])
def test_device_records_equal_host_records(L, case):
    data = cases.make_stream(case)
    host = capture_records(L, data, False)
    dev = capture_records(L, data, True)
    assert len(host) == len(dev) == case[3]
    for k, (h, d) in enumerate(zip(host, dev)):
        assert h[1] == d[1], "coefficient slots of picture %d differ" % k
        assert h[0] == d[0], "records of picture %d differ" % k


@pytest.mark.parametrize("lc", cases.LOSS, ids=[c[0] for c in cases.LOSS])
def test_device_records_equal_host_records_with_lost_slices(L, lc):
    data = cases.make_loss_stream(lc)
    host = capture_records(L, data, False)
    dev = capture_records(L, data, True)
    assert len(host) == len(dev)
    for k, (h, d) in enumerate(zip(host, dev)):
        assert h[0] == d[0], "records of picture %d differ" % k
        # the concealment list (4 bytes per H264B200_MB_CONCEAL macroblock) fills the last slots only partly
        n_conceal = sum(1 for i in range(len(h[0]) // 128) if h[0][i * 128] == 4)
        pad = (-(n_conceal * 4)) % 32
        assert len(h[1]) == len(d[1]) and h[1][:len(h[1]) - pad] == d[1][:len(d[1]) - pad], "slots of picture %d differ" % k


# --------------------------------------------------------------------------- the runner's device-parse pipeline
def run_streams(L, streams, threads, flags):
    n = len(streams)
    descs = (capi.StreamDesc * n)()
    for i, b in enumerate(streams):
        descs[i].data = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p); descs[i].len = len(b)
    res = [dict() for _ in streams]
    errs = [0] * n

    def _cb(user, stream, index, ptr, w, h, pic_id, err):
        res[stream][index] = hashlib.md5(ctypes.string_at(ptr, w * h * 3 // 2)).hexdigest()
        errs[stream] += err
    cb = capi.PICTURE_CB(_cb)
    rs = capi.RunStats()
    eng = L.h264b200EngineCreateEx(0, flags)
    try:
        rc = L.h264b200DecodeStreams(eng, descs, n, threads, ctypes.cast(cb, ctypes.c_void_p), None, ctypes.byref(rs))
    finally:
        L.h264b200EngineDestroy(eng)
    assert rc == 0 and rs.failed_streams == 0
    return [[d[i] for i in sorted(d)] for d in res], errs, rs


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_runner_device_parse_gives_the_golden_frames(L, golden, threads):
    sel = cases.SMALL[:12] + [c for c in cases.SMALL if c[0].startswith("dpb_")]
    md5s, errs, rs = run_streams(L, [cases.make_stream(c) for c in sel], threads, flags=1 | 8)
    for c, m in zip(sel, md5s):
        assert m == golden[c[0]]["frame_md5"], (c[0], threads)
    assert rs.pictures == sum(c[3] for c in sel) and sum(errs) == 0


def test_runner_device_parse_reports_concealed_macroblocks(L):
    sel = cases.LOSS
    md5s, errs, rs = run_streams(L, [cases.make_loss_stream(c) for c in sel], 4, flags=1 | 8)
    for c, m, e in zip(sel, md5s, errs):
        g = LOSS_GOLDEN[c[0]]
        assert m == g["frame_md5"], c[0]
        assert e == g["err_mbs"], c[0]


@pytest.mark.parametrize("threads,host", [(1, 1), (3, 5), (16, 7), (4, 99)])
def test_runner_host_share_of_a_device_parse_run(L, golden, threads, host, monkeypatch):
    """h264b200DecodeStreams gives `host` of the streams of a device-parse run to the worker threads' own parser
    (h264b200SetHostParse; spread evenly over the stream indices): host- and device-parsed instances share the engine and
    its rounds, and every picture equals the reference golden whoever parsed it (99: more than there are streams = all)."""
    monkeypatch.setenv("H264B200_HOST_STREAMS", str(host))
    sel = cases.SMALL[:12] + [c for c in cases.SMALL if c[0].startswith("dpb_")]
    md5s, errs, rs = run_streams(L, [cases.make_stream(c) for c in sel], threads, flags=1 | 8)
    for c, m in zip(sel, md5s):
        assert m == golden[c[0]]["frame_md5"], (c[0], threads, host)
    assert rs.pictures == sum(c[3] for c in sel) and sum(errs) == 0


def test_runner_host_share_reports_concealed_macroblocks(L, monkeypatch):
    monkeypatch.setenv("H264B200_HOST_STREAMS", str(len(cases.LOSS) // 2))
    sel = cases.LOSS
    md5s, errs, rs = run_streams(L, [cases.make_loss_stream(c) for c in sel], 4, flags=1 | 8)
    for c, m, e in zip(sel, md5s, errs):
        g = LOSS_GOLDEN[c[0]]
        assert m == g["frame_md5"], c[0]
        assert e == g["err_mbs"], c[0]
