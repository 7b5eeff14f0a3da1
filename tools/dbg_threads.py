import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from broadway_b200 import capi
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "streams.json")))
streams = [cases.make_stream(c) for c in cases.SMALL[:10]]
for rep in range(3):
    for threads in (1, 3, 10):
        with capi.Engine() as eng:
            md5s, rs = eng.decode_streams_md5(streams, threads=threads)
        bad = []
        for c, m in zip(cases.SMALL[:10], md5s):
            g = golden[c[0]]["frame_md5"]
            if m != g:
                bad.append((c[0], len(m), len(g), [i for i in range(min(len(m), len(g))) if m[i] != g[i]]))
        print("rep", rep, "threads", threads, "rounds", rs.rounds, "bad", bad)
