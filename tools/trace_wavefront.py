"""Debug: per-row wavefront timing of K4 for job 0 of a batch (H264B200_TRACE=n dumps n batches to stderr)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from broadway_b200 import bitstream, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
streams = [bitstream.synth(120, 68, 3, seed=1234 + i) for i in range(min(n, 8))]
streams = [streams[i % len(streams)] for i in range(n)]
with capi.Engine() as eng:
    eng.decode_streams(streams, threads=8)
