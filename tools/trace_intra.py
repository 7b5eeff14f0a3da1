"""Debug: K3/K4 wavefront timing on an all-intra 1080p picture (H264B200_TRACE=n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from broadway_b200 import bitstream, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s = bitstream.synth(120, 68, 2, seed=1234, intra_only=1)
with capi.Engine() as eng:
    eng.decode_streams([s] * n, threads=4)
