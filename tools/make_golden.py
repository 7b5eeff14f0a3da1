"""Generate tests/golden/streams.json: per-frame MD5 (Y|Cb|Cr, MB aligned) of every
case in tests/cases.py as decoded by the UNMODIFIED reference (oracle/_ref/refdec,
built by oracle/Makefile from /root/reference).  Run in the build container only;
the GPU box uses the committed JSON.  Also records the MD5 of each generated
stream so a change of the writer is detected instead of silently re-baselined."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases

def ref_md5(path):
    exe = os.path.join(ROOT, "oracle", "_ref", "refdec")
    r = subprocess.run([exe, "-m", path], capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    summary = json.loads(lines[-1])
    return [l.split()[2] for l in lines if l.startswith("frame ")], summary

out = {}
for case in cases.SMALL + cases.FULL:
    name, w, h, n, kw = case
    data = cases.make_stream(case)
    path = "/tmp/golden_%s.264" % name
    open(path, "wb").write(data)
    md5s, summary = ref_md5(path)
    assert len(md5s) == n and summary["err_mbs"] == 0, (name, len(md5s), summary)
    out[name] = {"width_mbs": w, "height_mbs": h, "frames": n, "overrides": kw,
                 "stream_md5": hashlib.md5(data).hexdigest(), "stream_bytes": len(data), "frame_md5": md5s}
    print(name, len(data), "bytes", n, "frames ok")
    os.remove(path)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "streams.json"), "w"), indent=1, sort_keys=True)

# packet-loss cases: the reference conceals; pin frames AND the concealed-macroblock count
def ref_md5_lossy(path):
    exe = os.path.join(ROOT, "oracle", "_ref", "refdec")
    r = subprocess.run([exe, "-m", path], capture_output=True, text=True)
    lines = r.stdout.splitlines()
    return [l.split()[2] for l in lines if l.startswith("frame ")], json.loads(lines[-1])

loss = {}
for lc in cases.LOSS:
    data = cases.make_loss_stream(lc)
    path = "/tmp/golden_%s.264" % lc[0]
    open(path, "wb").write(data)
    md5s, summary = ref_md5_lossy(path)
    loss[lc[0]] = {"stream_md5": hashlib.md5(data).hexdigest(), "frame_md5": md5s, "err_mbs": summary["err_mbs"]}
    print(lc[0], len(md5s), "frames,", summary["err_mbs"], "concealed macroblocks")
    os.remove(path)
json.dump(loss, open(os.path.join(ROOT, "tests", "golden", "loss.json"), "w"), indent=1, sort_keys=True)

# bench.py's parity gate: the first streams of its workload (rank 0), decoded by the reference
from broadway_b200 import bitstream
BENCH_FRAMES = 64          # bench.py DEFAULT_FRAMES
bench = {"width_mbs": 120, "height_mbs": 68, "frames": BENCH_FRAMES, "streams": []}
for i in range(2):
    data = bitstream.synth(120, 68, BENCH_FRAMES, seed=1234 + i)
    path = "/tmp/golden_bench_%d.264" % i
    open(path, "wb").write(data)
    md5s, summary = ref_md5(path)
    assert len(md5s) == BENCH_FRAMES and summary["err_mbs"] == 0
    bench["streams"].append({"seed": 1234 + i, "stream_md5": hashlib.md5(data).hexdigest(), "frame_md5": md5s})
    os.remove(path)
    print("bench stream", i, "ok")
json.dump(bench, open(os.path.join(ROOT, "tests", "golden", "bench_streams.json"), "w"), indent=1, sort_keys=True)
