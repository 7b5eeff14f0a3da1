"""Third-decoder cross-check (SURVEY.md section 4 item 3, section 8c "secondary check"; VERDICT r1 missing item 4):
the luma of every synthetic parity stream as FFmpeg's H.264 decoder reconstructs it (tests/golden/ffmpeg_luma.json,
made by tools/make_ffmpeg_golden.py) must equal what the oracle produces — so the in-repo writer and the reference
decoder cannot be agreeing on a non-conformant stream.  FFmpeg does not implement flexible macroblock ordering: the
seven FMO cases stay pinned by the reference alone (and say so in the fixture)."""
import hashlib
import json
import os
import subprocess
import tempfile

import pytest

import cases
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FF = json.load(open(os.path.join(ROOT, "tests", "golden", "ffmpeg_luma.json")))
CASES = [c for c in cases.SMALL + cases.FULL if "luma_md5" in FF[c[0]]]


def luma_md5s(frames, width, height, rows, cols):
    """MD5 of the top-left rows x cols of the luma plane of every I420 frame (FFmpeg applies the cropping rectangle,
    which the writer anchors at the top-left corner)."""
    import numpy as np
    fb = width * height * 3 // 2
    out = []
    for i in range(len(frames) // fb):
        y = np.frombuffer(frames, dtype=np.uint8, count=width * height, offset=i * fb).reshape(height, width)
        out.append(hashlib.md5(np.ascontiguousarray(y[:rows, :cols]).tobytes()).hexdigest())
    return out


def oracle_frames(data):
    with tempfile.TemporaryDirectory() as d:
        p, o = os.path.join(d, "s.264"), os.path.join(d, "s.yuv")
        open(p, "wb").write(data)
        r = subprocess.run([util.CPUCHK, "-o", o, p], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-300:]
        return open(o, "rb").read()


def test_fixture_covers_every_case():
    for c in cases.SMALL + cases.FULL:
        assert c[0] in FF
        assert ("luma_md5" in FF[c[0]]) != ("fmo_type" in c[4])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_luma_equals_ffmpeg(case):
    g = FF[case[0]]
    data = cases.make_stream(case)
    assert hashlib.md5(data).hexdigest() == g["stream_md5"], "the fixture was made from another stream: rerun tools/make_ffmpeg_golden.py"
    got = luma_md5s(oracle_frames(data), 16 * case[1], 16 * case[2], g["rows"], g["cols"])
    assert got == g["luma_md5"]


@pytest.mark.parametrize("case", [c for c in CASES if c in cases.SMALL][::4], ids=lambda c: c[0])
def test_live_ffmpeg_reproduces_fixture(case):
    """where OpenCV with an FFmpeg backend is installed, the fixture is reproducible"""
    pytest.importorskip("cv2")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ffmpeg_golden", os.path.join(ROOT, "tools", "make_ffmpeg_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    got = m.ffmpeg_luma_md5(cases.make_stream(case))
    if got is None:
        pytest.skip("this OpenCV build cannot decode H.264")
    assert [g[0] for g in got] == FF[case[0]]["luma_md5"]
