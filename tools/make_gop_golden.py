"""tests/golden/gop_4k.json: per-picture MD5 of the SEQUENTIAL decode, by the unmodified reference (oracle/_ref/refdec), of
the 4K streams bench.py --shard gop cuts into IDR-bounded segments and spreads over the GPUs.  Run here
(needs /root/reference for oracle/_ref):  python tools/make_gop_golden.py"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

bench._WL["name"] = "4k_gop"
streams = bench.gop_streams()
out = {"width_mbs": 240, "height_mbs": 135, "frames": bench.GOP_FRAMES, "idr_period": 4, "streams": []}
for i, data in enumerate(streams):
    path = "/tmp/gop_golden_%d.264" % i
    open(path, "wb").write(data)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "refdec"), "-m", path], capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    md5s = [l.split()[2] for l in lines if l.startswith("frame ")]
    assert len(md5s) == bench.GOP_FRAMES and json.loads(lines[-1])["err_mbs"] == 0
    out["streams"].append({"seed": 4321 + i, "stream_md5": hashlib.md5(data).hexdigest(), "stream_bytes": len(data), "frame_md5": md5s})
    os.remove(path)
    print("stream", i, len(data), "bytes ok")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "gop_4k.json"), "w"), indent=1, sort_keys=True)
