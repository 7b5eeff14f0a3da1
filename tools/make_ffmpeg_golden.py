"""Independent conformance check of the synthetic streams (SURVEY.md section 4 item 3 / section 8c "secondary check"):
decode every case of tests/cases.py that FFmpeg's H.264 decoder supports (everything except flexible macroblock
ordering) with OpenCV's FFmpeg backend and commit per-picture MD5s of the LUMA plane (cv2 hands out the luma plane only
with CAP_PROP_CONVERT_RGB = 0; FFmpeg applies the SPS cropping rectangle) to tests/golden/ffmpeg_luma.json.

Why: every parity stream comes from the in-repo writer, and the golden MD5s come from the reference decoder.  A third,
unrelated decoder agreeing on the luma of every stream rules out "writer and reference agree on a non-conformant
stream".  Run here (cv2 4.13 with FFmpeg is in the image):  python tools/make_ffmpeg_golden.py
"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ffmpeg_luma_md5(data):
    """[(md5 of the luma plane, rows, cols)] per output picture, or None when FFmpeg cannot decode the stream."""
    import cv2
    import numpy as np
    with tempfile.NamedTemporaryFile(suffix=".264", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
        out = []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            fr = np.ascontiguousarray(fr)
            if fr.ndim != 2:
                return None
            out.append((hashlib.md5(fr.tobytes()).hexdigest(), int(fr.shape[0]), int(fr.shape[1])))
        cap.release()
    finally:
        os.remove(path)
    return out or None


def main():
    import cases
    import __graft_entry__
    __graft_entry__.build()
    res = {}
    for case in cases.SMALL + cases.FULL:
        name = case[0]
        if "fmo_type" in case[4]:
            res[name] = {"unsupported": "FFmpeg's H.264 decoder does not implement flexible macroblock ordering"}
            continue
        data = cases.make_stream(case)
        got = ffmpeg_luma_md5(data)
        assert got and len(got) == case[3], (name, got and len(got))
        res[name] = {"stream_md5": hashlib.md5(data).hexdigest(), "rows": got[0][1], "cols": got[0][2], "luma_md5": [g[0] for g in got]}
        print(name, len(got), "pictures", got[0][1:], file=sys.stderr)
    p = os.path.join(ROOT, "tests", "golden", "ffmpeg_luma.json")
    json.dump(res, open(p, "w"), indent=0, sort_keys=True)
    print("wrote", p)


if __name__ == "__main__":
    main()
