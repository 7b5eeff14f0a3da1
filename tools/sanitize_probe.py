"""Small decode workload for compute-sanitizer (memcheck / racecheck): a few tiny streams covering every kernel
(K1..K5, k3c) through the single-stream API and the batched runner."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import json
import cases
from broadway_b200 import capi
golden = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "streams.json")))
names = ["one_mb", "odd_size", "far_mv", "p_intra_mix", "multi_slice", "fmo_box_out", "dpb_long_term"]
sel = [c for c in cases.SMALL if c[0] in names]
for c in sel:
    got, info = capi.decode_annexb(cases.make_stream(c))
    assert got == golden[c[0]]["frame_md5"], c[0]
with capi.Engine() as eng:
    md5s, rs = eng.decode_streams_md5([cases.make_stream(c) for c in sel] + [cases.make_loss_stream(cases.LOSS[4])], threads=2)
    for c, m in zip(sel, md5s):
        assert m == golden[c[0]]["frame_md5"], c[0]
print("sanitize probe ok:", len(sel), "streams")
