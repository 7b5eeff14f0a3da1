"""Helpers of the test-suite.  The ORACLE side lives here: oracle/cpuchkdec (the CPU
restatement under the product's host parser) and oracle/_ref/refdec (the unmodified
reference, present only where it was built).  Tests are the only place besides
bench.py's cpu_baseline / reference arm and __graft_entry__.smoke() that may execute
anything under oracle/."""
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPUCHK = os.path.join(ROOT, "oracle", "cpuchkdec")
REFDEC = os.path.join(ROOT, "oracle", "_ref", "refdec")


def _run_cli(exe, data, extra=()):
    with tempfile.NamedTemporaryFile(suffix=".264", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        r = subprocess.run([exe, "-m", *extra, path], capture_output=True, text=True, timeout=600)
    finally:
        os.remove(path)
    lines = r.stdout.splitlines()
    # exit code 1 with a summary line = "decoded, but macroblocks were concealed" (like DecTestBench.c:424-428)
    assert r.returncode == 0 or (r.returncode == 1 and lines and lines[-1].startswith("{")), (exe, r.returncode, r.stderr[-400:])
    return [l.split()[2] for l in lines if l.startswith("frame ")], json.loads(lines[-1])


def oracle_md5(data):
    """Per-frame MD5 from the CPU restatement (records -> pixels on the CPU)."""
    return _run_cli(CPUCHK, data)


def reference_md5(data):
    """Per-frame MD5 from the unmodified reference build, or None where it is absent."""
    if not os.path.exists(REFDEC):
        return None
    return _run_cli(REFDEC, data)


# ----------------------------------------------------------------------------- records of the host parser (oracle side)
import ctypes  # noqa: E402

CPUCHK_LIB = os.path.join(ROOT, "oracle", "libh264b200_cpuchk.so")
TAP_RECORDS = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32)


class Tap(ctypes.Structure):
    _fields_ = [("user", ctypes.c_void_p), ("records", TAP_RECORDS), ("residual", ctypes.c_void_p), ("predeblock", ctypes.c_void_p)]


_cpuchk = None


def cpuchk_lib():
    """oracle/libh264b200_cpuchk.so: the product's host decoder + runner over the CPU restatement backend (test-only)."""
    global _cpuchk
    if _cpuchk is not None:
        return _cpuchk
    from broadway_b200 import capi
    lib = ctypes.CDLL(CPUCHK_LIB)
    vp, u32 = ctypes.c_void_p, ctypes.c_uint32
    sp = ctypes.POINTER(capi.Storage)
    lib.h264b200EngineCreateEx.argtypes = [ctypes.c_int, u32]; lib.h264b200EngineCreateEx.restype = vp
    lib.h264b200EngineDestroy.argtypes = [vp]; lib.h264b200EngineDestroy.restype = None
    lib.h264b200InitOnEngine.argtypes = [sp, u32, vp]; lib.h264b200InitOnEngine.restype = u32
    lib.h264bsdDecode.argtypes = [sp, vp, u32, u32, ctypes.POINTER(u32)]; lib.h264bsdDecode.restype = u32
    lib.h264bsdFlushBuffer.argtypes = [sp]; lib.h264bsdFlushBuffer.restype = None
    lib.h264bsdShutdown.argtypes = [sp]; lib.h264bsdShutdown.restype = None
    lib.recon_cpu_set_tap.argtypes = [ctypes.POINTER(Tap)]; lib.recon_cpu_set_tap.restype = None
    lib.h264b200DecodeStreams.argtypes = [vp, ctypes.POINTER(capi.StreamDesc), u32, u32, vp, vp, ctypes.POINTER(capi.RunStats)]
    lib.h264b200DecodeStreams.restype = ctypes.c_int
    _cpuchk = lib
    return lib


def canonical_records(mbs, n_mbs):
    """Bytes of the records with the one field no kernel reads blanked: the vectors of macroblocks that are not inter
    (the host parser leaves whatever the buffer held there)."""
    out = bytearray(mbs)
    for i in range(n_mbs):
        r = i * 128
        if out[r] != 0:                                  # not H264B200_MB_INTER: mv[] is never read
            out[r + 64:r + 128] = bytes(64)
    return bytes(out)


def capture_records(L, data, device_parse):
    """Decode `data` on the CPU backend of `L` (cpuchk_lib()) and return [(records, slots)] per picture, in decoding
    order: the host parser's output (device_parse False) or kp_core.h's compiled for the CPU (True)."""
    from broadway_b200 import capi
    pics = []

    def on_records(user, mbs, n_mbs, coef, n_slots):
        pics.append((canonical_records(ctypes.string_at(mbs, n_mbs * 128), n_mbs), ctypes.string_at(coef, n_slots * 32)))
    cb = TAP_RECORDS(on_records)
    tap = Tap(None, cb, None, None)
    L.recon_cpu_set_tap(ctypes.byref(tap))
    eng = L.h264b200EngineCreateEx(0, 8 if device_parse else 0)
    st = capi.Storage()
    try:
        assert L.h264b200InitOnEngine(ctypes.byref(st), 0, eng) == 0
        buf = ctypes.create_string_buffer(bytes(data), len(data) + 16)
        base, pos, n = ctypes.addressof(buf), 0, len(data)
        nread = ctypes.c_uint32()
        while pos < n:
            rc = L.h264bsdDecode(ctypes.byref(st), base + pos, n - pos, 0, ctypes.byref(nread))
            pos += nread.value
            if nread.value == 0 and rc not in (capi.H264BSD_PIC_RDY, capi.H264BSD_HDRS_RDY):
                break
        L.h264bsdFlushBuffer(ctypes.byref(st))
    finally:
        L.h264bsdShutdown(ctypes.byref(st))
        L.h264b200EngineDestroy(eng)
        L.recon_cpu_set_tap(None)
    return pics
