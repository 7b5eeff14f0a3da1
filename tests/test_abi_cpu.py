"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports
every symbol include/*.h declares, refuses to reconstruct without CUDA (no CPU
fallback), and the GOP splitter (host logic) is exact."""
import ctypes
import os
import re

import pytest

import cases
import util
from broadway_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("h264b200.h", "h264b200_swdec.h", "h264b200_batch.h", "h264b200_shim.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b((?:h264bsd|H264SwDec|h264b200|broadway(?=[A-Z]))[A-Za-z0-9_]*)\s*\(", src):
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(capi.LIB_PATH)
    decl = _declared_symbols()
    assert {"h264bsdInit", "h264bsdDecode", "h264bsdNextOutputPicture", "h264bsdShutdown", "H264SwDecDecode",
            "h264b200DecodeStreams", "h264b200SplitGops"} <= decl
    missing = [n for n in sorted(decl) if not hasattr(L, n)]
    assert not missing, missing


def test_writer_library_exports():
    L = ctypes.CDLL(os.path.join(ROOT, "broadway_b200", "libh264writer.so"))
    for n in ("h264w_default_params", "h264w_bound", "h264w_generate"):
        assert hasattr(L, n)


def test_no_cpu_fallback_without_cuda():
    """Without a usable CUDA device the product must FAIL (H264BSD_MEMALLOC_ERROR), not reconstruct on the CPU."""
    rc, msg = capi.probe()
    if rc == 0:
        pytest.skip("a CUDA device is present: " + msg)
    data = cases.make_stream(cases.SMALL[0])
    with pytest.raises(RuntimeError):
        capi.decode_annexb(data)
    L = capi.lib()
    st = capi.Storage()
    assert L.h264bsdInit(ctypes.byref(st), 0) == 0
    buf = ctypes.create_string_buffer(data, len(data) + 16)
    pos, nread, seen = 0, ctypes.c_uint32(), set()
    while pos < len(data):
        rc = L.h264bsdDecode(ctypes.byref(st), ctypes.addressof(buf) + pos, len(data) - pos, 0, ctypes.byref(nread))
        seen.add(rc)
        if rc == capi.H264BSD_MEMALLOC_ERROR or (nread.value == 0 and rc != capi.H264BSD_HDRS_RDY):
            break
        pos += nread.value
    L.h264bsdShutdown(ctypes.byref(st))
    assert capi.H264BSD_MEMALLOC_ERROR in seen and capi.H264BSD_PIC_RDY not in seen


def test_split_gops_segments_decode_like_the_whole_stream(golden):
    """IDR-bounded segments are self-contained: decoding them one by one (oracle) gives the
    frames of the sequential decode, in order."""
    case = next(c for c in cases.SMALL if c[0] == "idr_period")
    data = cases.make_stream(case)
    segs = capi.split_gops(data)
    assert len(segs) == 3
    got = []
    for s in segs:
        md5s, summary = util.oracle_md5(s)
        assert summary["err_mbs"] == 0 and len(md5s) == 3
        got += md5s
    assert got == golden["idr_period"]["frame_md5"]


def test_split_gops_single_idr_and_empty():
    data = cases.make_stream(cases.SMALL[0])
    segs = capi.split_gops(data)
    assert len(segs) == 1 and segs[0] == data
    assert capi.split_gops(b"\x00\x00\x00\x01\x09\x10") == []


HOOK_PROBE = r"""
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "h264b200_swdec.h"
/* an embedder's own hooks, like every test bench of the reference defines them (DecTestBench.c:678-760) */
static unsigned n_malloc, n_free, n_memset; static long live;
void *H264SwDecMalloc(u32 size) { n_malloc++; live++; return malloc(size); }
void H264SwDecFree(void *p) { n_free++; live--; free(p); }
void H264SwDecMemset(void *p, i32 v, u32 n) { n_memset++; memset(p, v, n); }
void H264SwDecMemcpy(void *d, void *s, u32 n) { memcpy(d, s, n); }
void H264SwDecTrace(char *s) { (void)s; }
int main(int argc, char **argv)
{
    H264SwDecInst inst; H264SwDecInput in; H264SwDecOutput out;
    FILE *f = fopen(argv[1], "rb"); static unsigned char buf[1 << 20]; size_t n = fread(buf, 1, sizeof buf - 64, f); fclose(f);
    if (H264SwDecInit(&inst, 0) != H264SWDEC_OK) return 2;
    in.pStream = buf; in.dataLen = (u32)n; in.picId = 0; in.intraConcealmentMethod = 0;
    H264SwDecDecode(inst, &in, &out);            /* parameter sets are stored (host heap) whether or not a GPU is present */
    H264SwDecRelease(inst);
    printf("%u %u %u %ld\n", n_malloc, n_free, n_memset, live);
    return 0;
}
"""


def test_swdec_memory_hooks_are_honoured(tmp_path):
    """H264SwDecApi.h:158-173: the embedder's H264SwDecMalloc / Free / Memset replace the library's weak defaults, and
    every host-side allocation of the decoder goes through them (none leaks)."""
    import subprocess
    src = tmp_path / "hooks.c"
    src.write_text(HOOK_PROBE)
    exe = tmp_path / "hooks"
    pkg = os.path.join(ROOT, "broadway_b200")
    subprocess.run(["gcc", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L" + pkg, "-lh264b200",
                    "-Wl,-rpath," + pkg], check=True)
    stream = tmp_path / "s.264"
    data = cases.make_stream(cases.SMALL[0])
    stream.write_bytes(data[:data.index(b"\x00\x00\x00\x01", data.index(b"\x00\x00\x00\x01", 8) + 4)])    # SPS + PPS only
    r = subprocess.run([str(exe), str(stream)], capture_output=True, text=True, check=True)
    n_malloc, n_free, n_memset, live = map(int, r.stdout.split())
    assert n_malloc >= 3 and n_memset >= 1            # container, decoder state, parameter sets
    assert n_malloc == n_free and live == 0
