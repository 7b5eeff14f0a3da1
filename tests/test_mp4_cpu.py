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
