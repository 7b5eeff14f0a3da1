"""CPU tests of the MP4 -> Annex-B demuxer (C counterpart of Player/mp4.js): muxing a synthetic stream
into an MP4 and demuxing it back yields the same NAL sequence, and therefore (oracle) the same frames."""
import pytest

import cases
import mp4mux
import util
from broadway_b200 import capi


@pytest.mark.parametrize("kw", [dict(), dict(samples_per_chunk=1), dict(samples_per_chunk=100), dict(length_size=2), dict(use_co64=True)],
                         ids=["default", "chunk1", "one_chunk", "len2", "co64"])
def test_mux_demux_round_trip(kw, golden):
    case = next(c for c in cases.SMALL if c[0] == "multi_slice")
    data = cases.make_stream(case)
    mp4, n_samples = mp4mux.mux(data, 16 * case[1], 16 * case[2], **kw)
    assert n_samples == case[3]
    annexb, n = capi.mp4_to_annexb(mp4)
    assert n == n_samples
    assert mp4mux.split_annexb(annexb) == mp4mux.split_annexb(data)
    md5s, summary = util.oracle_md5(annexb)
    assert md5s == golden["multi_slice"]["frame_md5"]


def test_rejects_garbage_and_truncation():
    case = cases.SMALL[0]
    mp4, _ = mp4mux.mux(cases.make_stream(case), 320, 192)
    with pytest.raises(ValueError):
        capi.mp4_to_annexb(b"not an mp4 file at all")
    with pytest.raises(ValueError):
        capi.mp4_to_annexb(mp4[:len(mp4) // 2])
    with pytest.raises(ValueError):
        capi.mp4_to_annexb(mp4.replace(b"avc1", b"hvc1"))


def _box64(kind, payload, size):
    """box with a 64-bit largesize field that claims `size` bytes"""
    import struct
    return struct.pack(">I4sQ", 1, kind, size) + payload


def test_rejects_wrapping_box_sizes_and_offsets():
    """Crafted sizes / chunk offsets near 2^64 must be rejected, not wrapped past the bounds checks
    (a matching box would otherwise span ~2^64 bytes; a mismatching one would move the scan backwards for ever)."""
    import struct
    case = cases.SMALL[0]
    data = cases.make_stream(case)
    mp4, _ = mp4mux.mux(data, 320, 192, use_co64=True)
    pad = mp4mux.box(b"free", b"\0" * 16)
    for kind in (b"moov", b"skip"):                          # type match and type mismatch
        for size in (2 ** 64 - 16, 2 ** 64 - 1, 2 ** 63):
            with pytest.raises(ValueError):
                capi.mp4_to_annexb(pad + _box64(kind, b"\0" * 64, size))
    # the same inside moov (the trak loop has its own copy of the scan)
    inner = _box64(b"trak", b"\0" * 32, 2 ** 64 - 24)
    with pytest.raises(ValueError):
        capi.mp4_to_annexb(mp4mux.box(b"moov", mp4mux.box(b"mvhd", b"\0" * 8) + inner))
    # co64 chunk offsets close to 2^64: pos + sample_size wraps
    at = mp4.index(b"co64")
    n_chunks = struct.unpack(">I", mp4[at + 8:at + 12])[0]
    assert n_chunks >= 1
    for off in (2 ** 64 - 1, 2 ** 64 - 8, 2 ** 64 - 100, len(mp4) + 1):
        bad = bytearray(mp4)
        bad[at + 12:at + 20] = struct.pack(">Q", off)
        with pytest.raises(ValueError):
            capi.mp4_to_annexb(bytes(bad))
    # NAL length prefix larger than its sample
    annexb, n = capi.mp4_to_annexb(mp4)
    first = struct.unpack(">Q", mp4[at + 12:at + 20])[0]
    bad = bytearray(mp4)
    bad[first:first + 4] = b"\xff\xff\xff\xf0"
    with pytest.raises(ValueError):
        capi.mp4_to_annexb(bytes(bad))
