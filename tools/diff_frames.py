"""Debug aid: decode a small synthetic stream with the CUDA build (b200dec) and with the CPU restatement
(oracle/cpuchkdec), and print WHERE the first differing frame differs (per macroblock and per position
inside the macroblock).  usage: python tools/diff_frames.py [w_mbs h_mbs n_frames key=value ...]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from broadway_b200 import bitstream as bs

w, h, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (6, 4, 1)
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[4:])}
data = bs.synth(w, h, n, seed=7, **kw)
open("/tmp/dbg.264", "wb").write(data)
outs = {}
for name, exe in (("gpu", "broadway_b200/bin/b200dec"), ("cpu", "oracle/cpuchkdec")):
    out = "/tmp/dbg_%s.yuv" % name
    if os.path.exists(out):
        os.remove(out)
    subprocess.run([os.path.join(ROOT, exe), "-o", out, "/tmp/dbg.264"], capture_output=True, timeout=120)
    outs[name] = np.fromfile(out, dtype=np.uint8)
W, H = 16 * w, 16 * h
fb = W * H * 3 // 2
print("bytes", {k: v.size for k, v in outs.items()}, "frame", fb)
for f in range(min(outs["gpu"].size, outs["cpu"].size) // fb):
    a = outs["gpu"][f * fb:(f + 1) * fb]; b = outs["cpu"][f * fb:(f + 1) * fb]
    if (a == b).all():
        print("frame", f, "identical"); continue
    print("frame", f, "differs")
    for pname, off, pw, ph, mb in (("Y", 0, W, H, 16), ("Cb", W * H, W // 2, H // 2, 8), ("Cr", W * H * 5 // 4, W // 2, H // 2, 8)):
        pa = a[off:off + pw * ph].reshape(ph, pw).astype(int); pb = b[off:off + pw * ph].reshape(ph, pw).astype(int)
        d = pa != pb
        if not d.any():
            print(" ", pname, "ok"); continue
        print(" ", pname, "differing samples:", int(d.sum()))
        grid = d.reshape(ph // mb, mb, pw // mb, mb).sum(axis=(1, 3))
        print("   per macroblock:\n" + "\n".join("    " + " ".join("%3d" % v for v in row) for row in grid))
        print("   by row inside MB:", d.reshape(ph // mb, mb, pw).sum(axis=(0, 2)).tolist())
        print("   by col inside MB:", d.reshape(ph, pw // mb, mb).sum(axis=(0, 1)).tolist())
        ys, xs = np.nonzero(d)
        for y, x in list(zip(ys, xs))[:12]:
            print("     (x %d y %d) gpu %d cpu %d" % (x, y, pa[y, x], pb[y, x]))
    break
