"""The multi-stream runner (broadway_b200/csrc/h264_runner.c, h264b200DecodeStreams) on the CPU.

oracle/libh264b200_cpuchk.so links the PRODUCT's runner source over the CPU restatement backend through
oracle/engine_shim.c, so the host-side scheduling — (round, stream) work items claimed by any thread, two launch
groups, hand-over of a decoder between threads, output collected one round late, streams of different lengths
finishing at different rounds — is exercised without a GPU.  Every picture must equal the reference golden,
whatever the thread count and grouping (the first dynamic version of the runner lost whole streams at
threads == streams; that case is pinned here)."""
import ctypes
import hashlib
import json
import os

import pytest

import cases
from broadway_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libh264b200_cpuchk.so")


@pytest.fixture(scope="module")
def L():
    lib = ctypes.CDLL(LIB)
    vp = ctypes.c_void_p
    lib.h264b200EngineCreateEx.argtypes = [ctypes.c_int, ctypes.c_uint32]; lib.h264b200EngineCreateEx.restype = vp
    lib.h264b200EngineDestroy.argtypes = [vp]; lib.h264b200EngineDestroy.restype = None
    lib.h264b200DecodeStreams.argtypes = [vp, ctypes.POINTER(capi.StreamDesc), ctypes.c_uint32, ctypes.c_uint32, vp, vp, ctypes.POINTER(capi.RunStats)]
    lib.h264b200DecodeStreams.restype = ctypes.c_int
    return lib


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "streams.json")))


def run(L, streams, threads, flags=1):
    n = len(streams)
    descs = (capi.StreamDesc * n)()
    for i, b in enumerate(streams):
        descs[i].data = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p); descs[i].len = len(b)
    res = [dict() for _ in streams]

    def _cb(user, stream, index, ptr, w, h, pic_id, err):
        res[stream][index] = hashlib.md5(ctypes.string_at(ptr, w * h * 3 // 2)).hexdigest()
    cb = capi.PICTURE_CB(_cb)
    rs = capi.RunStats()
    eng = L.h264b200EngineCreateEx(0, flags)
    try:
        rc = L.h264b200DecodeStreams(eng, descs, n, threads, ctypes.cast(cb, ctypes.c_void_p), None, ctypes.byref(rs))
    finally:
        L.h264b200EngineDestroy(eng)
    assert rc == 0 and rs.failed_streams == 0
    return [[d[i] for i in sorted(d)] for d in res], rs


@pytest.mark.parametrize("threads", [1, 2, 3, 10, 16])
def test_every_thread_count_gives_the_golden_frames(L, golden, threads):
    sel = cases.SMALL[:10]                              # 2 .. 8 pictures per stream: streams finish at different rounds
    md5s, rs = run(L, [cases.make_stream(c) for c in sel], threads)
    for c, m in zip(sel, md5s):
        assert m == golden[c[0]]["frame_md5"], (c[0], threads)
    assert rs.pictures == sum(c[3] for c in sel)
    assert rs.threads == min(threads, len(sel))


def test_two_launch_groups_and_whole_round_batches(L, golden):
    """>= 4 streams per thread: two groups per round (twice the batches); with RETAIN set a batch is a whole round."""
    sel = cases.SMALL[:12]
    streams = [cases.make_stream(c) for c in sel]
    longest = max(c[3] for c in sel)
    md5s, rs = run(L, streams, threads=2, flags=1)
    for c, m in zip(sel, md5s):
        assert m == golden[c[0]]["frame_md5"], c[0]
    assert rs.rounds >= 2 * longest
    md5s1, rs1 = run(L, streams, threads=2, flags=1 | 2)
    assert md5s1 == md5s
    assert longest <= rs1.rounds <= longest + 2


def test_many_repeats_of_one_stream(L, golden):
    """The same stream 24 times on 5 threads: decoder instances are independent, results identical."""
    c = next(x for x in cases.SMALL if x[0] == "p_intra_mix")
    md5s, _ = run(L, [cases.make_stream(c)] * 24, threads=5)
    for m in md5s:
        assert m == golden[c[0]]["frame_md5"]


EPB_HEAVY = [dict(max_level=2000, qp=0, qp_jitter=0, max_coeffs=4, coded_blk_permille=900),
             dict(max_level=1000, qp=2, qp_jitter=0, max_coeffs=2, coded_blk_permille=950, slices_per_pic=3),
             dict(first_idr_ipcm=1, max_level=300, qp=1, qp_jitter=0, max_coeffs=3, coded_blk_permille=800)]


@pytest.mark.parametrize("flags", [1, 1 | 8], ids=["host_parse", "device_parse"])
def test_read_only_input_with_emulation_prevention_bytes(L, flags):
    """h264b200DecodeStreams decodes straight from the caller's bytes (h264b200SetReadOnlyInput): NAL units that
    contain emulation prevention bytes (dozens per stream here) are unescaped into decoder-owned memory.  The caller's
    streams must come back untouched — several instances share one copy — and the frames must equal those of the
    in-place decode (the oracle CLI, h264bsd_byte_stream.c:192-234 semantics) and of the reference where it is built."""
    import util
    from broadway_b200 import bitstream
    streams = [bitstream.synth(20, 12, 5, seed=99, **kw) for kw in EPB_HEAVY]
    assert all(s.count(b"\x00\x00\x03") >= 9 for s in streams)
    shared = [streams[0], streams[1], streams[2], streams[0], streams[0], streams[1]]     # the same bytes objects, concurrently
    before = [hashlib.md5(s).hexdigest() for s in streams]
    md5s, _ = run(L, shared, threads=3, flags=flags)
    assert [hashlib.md5(s).hexdigest() for s in streams] == before, "the runner wrote to the caller's streams"
    want = [util.oracle_md5(s)[0] for s in streams]
    assert md5s == [want[0], want[1], want[2], want[0], want[0], want[1]]
    for s, w in zip(streams, want):
        ref = util.reference_md5(s)
        if ref is not None:
            assert ref[0] == w
