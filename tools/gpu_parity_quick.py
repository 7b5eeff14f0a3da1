"""Quick GPU parity probe: b200dec (CUDA) vs refdec (reference, CPU) vs cpuchkdec (restatement) per-frame MD5."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from broadway_b200 import bitstream as bs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    ("ipcm_p16", dict(w=20, h=12, n=4, kw=dict(first_idr_ipcm=1, part_mix=0, coded_blk_permille=0, p_intra_permille=0, p_skip_permille=0, deblock_idc=1))),
    ("ipcm_p16_dbk", dict(w=20, h=12, n=4, kw=dict(first_idr_ipcm=1, part_mix=0, coded_blk_permille=0, p_intra_permille=0, p_skip_permille=0))),
    ("ipcm_pmix_res", dict(w=20, h=12, n=4, kw=dict(first_idr_ipcm=1, p_intra_permille=0, deblock_idc=1))),
    ("intra_nodbk", dict(w=20, h=12, n=3, kw=dict(intra_only=1, deblock_idc=1))),
    ("intra", dict(w=20, h=12, n=3, kw=dict(intra_only=1))),
    ("ippp", dict(w=20, h=12, n=8, kw=dict())),
    ("mix", dict(w=20, h=12, n=8, kw=dict(p_intra_permille=100, num_ref_frames=3, slices_per_pic=3, multi_slice_params=1, qp_jitter=6))),
    ("p1080", dict(w=120, h=68, n=5, kw=dict())),
]
def md5s(exe, path):
    r = subprocess.run([exe, "-m", path], capture_output=True, text=True, timeout=120)
    return [l for l in r.stdout.splitlines() if l.startswith("frame")], r.stdout.splitlines()[-1:] , r.stderr[-500:]
bad = 0
for name, c in CASES:
    data = bs.synth(c["w"], c["h"], c["n"], seed=7, **c["kw"])
    path = f"/tmp/{name}.264"
    open(path, "wb").write(data)
    ref, _, _ = md5s(os.path.join(ROOT, "oracle/_ref/refdec"), path)
    try:
        gpu, tail, err = md5s(os.path.join(ROOT, "broadway_b200/bin/b200dec"), path)
    except subprocess.TimeoutExpired:
        print(name, "TIMEOUT"); bad += 1; continue
    ok = ref == gpu and len(ref) == c["n"]
    first_bad = next((i for i, (a, b) in enumerate(zip(ref, gpu)) if a != b), None)
    print(name, "OK" if ok else "MISMATCH", len(ref), len(gpu), "first_bad", first_bad, tail, err if not ok else "")
    bad += not ok
sys.exit(1 if bad else 0)
