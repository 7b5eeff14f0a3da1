"""ctypes binding of libh264b200.so for tests and bench.py.

This is NOT the product's host side (that is C: csrc/h264_decoder.c, h264_slice.c,
h264_runner.c ... behind the reference's own entry points); it only lets Python
drive the C ABI declared in include/h264b200.h, h264b200_swdec.h and
h264b200_batch.h.  It never falls back to anything: a missing library or an
unusable CUDA device raises.
"""
import ctypes
import hashlib
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("H264B200_LIB") or os.path.join(_HERE, "libh264b200.so")   # the override is for A/B builds of the host parser

H264BSD_RDY, H264BSD_PIC_RDY, H264BSD_HDRS_RDY, H264BSD_ERROR, H264BSD_PARAM_SET_ERROR, H264BSD_MEMALLOC_ERROR = range(6)
H264SWDEC_OK, H264SWDEC_STRM_PROCESSED, H264SWDEC_PIC_RDY, H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY, H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY = range(5)
ENGINE_BATCHED, ENGINE_RETAIN, ENGINE_NO_D2H, ENGINE_DEVICE_PARSE, ENGINE_NO_RECON, ENGINE_TAP_PREDEBLOCK = 1, 2, 4, 8, 16, 32


class Storage(ctypes.Structure):
    _fields_ = [("impl", ctypes.c_void_p), ("reserved", ctypes.c_uint64 * 7)]


class SwDecInput(ctypes.Structure):
    _fields_ = [("pStream", ctypes.c_void_p), ("dataLen", ctypes.c_uint32), ("picId", ctypes.c_uint32),
                ("intraConcealmentMethod", ctypes.c_uint32)]


class SwDecOutput(ctypes.Structure):
    _fields_ = [("pStrmCurrPos", ctypes.c_void_p)]


class SwDecPicture(ctypes.Structure):
    _fields_ = [("pOutputPicture", ctypes.c_void_p), ("picId", ctypes.c_uint32), ("isIdrPicture", ctypes.c_uint32),
                ("nbrOfErrMBs", ctypes.c_uint32)]


class SwDecInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("profile", "picWidth", "picHeight", "videoRange", "matrixCoefficients",
                                               "parWidth", "parHeight", "croppingFlag", "cropLeftOffset", "cropOutWidth",
                                               "cropTopOffset", "cropOutHeight")]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("kernel_launches", "pictures", "h2d_bytes", "d2h_bytes", "batches", "kp_launches", "kp_pictures")]


class KernelTimes(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * 5), ("bytes", ctypes.c_uint64 * 5), ("launches", ctypes.c_uint64 * 5)]


class StreamDesc(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("len", ctypes.c_size_t)]


class RunStats(ctypes.Structure):
    _fields_ = [("pictures", ctypes.c_uint64), ("bytes_in", ctypes.c_uint64), ("bytes_out", ctypes.c_uint64),
                ("err_mbs", ctypes.c_uint32), ("failed_streams", ctypes.c_uint32), ("rounds", ctypes.c_uint32),
                ("threads", ctypes.c_uint32), ("seconds", ctypes.c_double), ("parse_seconds", ctypes.c_double),
                ("wait_seconds", ctypes.c_double), ("host_streams", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


PICTURE_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                              ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32)

_lib = None


def lib():
    """Load libh264b200.so (no compute is started by loading it)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u32p = ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)
    sp = ctypes.POINTER(Storage)
    L.h264bsdInit.argtypes = [sp, u32]; L.h264bsdInit.restype = u32
    L.h264bsdDecode.argtypes = [sp, vp, u32, u32, u32p]; L.h264bsdDecode.restype = u32
    L.h264bsdNextOutputPicture.argtypes = [sp, u32p, u32p, u32p]; L.h264bsdNextOutputPicture.restype = vp
    L.h264bsdShutdown.argtypes = [sp]; L.h264bsdShutdown.restype = None
    L.h264bsdPicWidth.argtypes = [sp]; L.h264bsdPicWidth.restype = u32
    L.h264bsdPicHeight.argtypes = [sp]; L.h264bsdPicHeight.restype = u32
    L.h264bsdFlushBuffer.argtypes = [sp]; L.h264bsdFlushBuffer.restype = None
    L.h264bsdCroppingParams.argtypes = [sp, u32p, u32p, u32p, u32p, u32p]; L.h264bsdCroppingParams.restype = None
    L.H264SwDecInit.argtypes = [ctypes.POINTER(vp), u32]; L.H264SwDecInit.restype = ctypes.c_int
    L.H264SwDecDecode.argtypes = [vp, ctypes.POINTER(SwDecInput), ctypes.POINTER(SwDecOutput)]; L.H264SwDecDecode.restype = ctypes.c_int
    L.H264SwDecNextPicture.argtypes = [vp, ctypes.POINTER(SwDecPicture), u32]; L.H264SwDecNextPicture.restype = ctypes.c_int
    L.H264SwDecGetInfo.argtypes = [vp, ctypes.POINTER(SwDecInfo)]; L.H264SwDecGetInfo.restype = ctypes.c_int
    L.H264SwDecRelease.argtypes = [vp]; L.H264SwDecRelease.restype = None
    L.h264b200Probe.argtypes = [ctypes.c_char_p, ctypes.c_size_t]; L.h264b200Probe.restype = ctypes.c_int
    L.h264b200EngineCreate.argtypes = [ctypes.c_int]; L.h264b200EngineCreate.restype = vp
    L.h264b200EngineCreateEx.argtypes = [ctypes.c_int, u32]; L.h264b200EngineCreateEx.restype = vp
    L.h264b200EngineDestroy.argtypes = [vp]; L.h264b200EngineDestroy.restype = None
    L.h264b200EngineSetFlags.argtypes = [vp, u32]; L.h264b200EngineSetFlags.restype = None
    L.h264b200InitOnEngine.argtypes = [sp, u32, vp]; L.h264b200InitOnEngine.restype = u32
    L.h264b200EngineSubmit.argtypes = [vp]; L.h264b200EngineSubmit.restype = u32
    L.h264b200EngineSync.argtypes = [vp]; L.h264b200EngineSync.restype = None
    L.h264b200EngineStats.argtypes = [vp, ctypes.POINTER(Stats)]; L.h264b200EngineStats.restype = None
    L.h264b200EngineErrorFlags.argtypes = [vp]; L.h264b200EngineErrorFlags.restype = u32
    L.h264b200EngineReplay.argtypes = [vp, u32, ctypes.c_int]; L.h264b200EngineReplay.restype = u32
    L.h264b200EngineReplayMs.argtypes = [vp]; L.h264b200EngineReplayMs.restype = ctypes.c_double
    L.h264b200EngineDropRetained.argtypes = [vp]; L.h264b200EngineDropRetained.restype = None
    L.h264b200EngineCheckResident.argtypes = [vp]; L.h264b200EngineCheckResident.restype = u32
    L.h264b200EngineKernelTimes.argtypes = [vp, ctypes.POINTER(KernelTimes), ctypes.c_int]; L.h264b200EngineKernelTimes.restype = None
    L.h264b200NextOutputPictureAsync.argtypes = [sp, u32p, u32p, u32p, u32p]; L.h264b200NextOutputPictureAsync.restype = vp
    L.h264b200PictureWait.argtypes = [sp, u32]; L.h264b200PictureWait.restype = u32
    L.h264b200EngineAdvance.argtypes = [vp]; L.h264b200EngineAdvance.restype = u32
    L.h264b200EngineSetWindow.argtypes = [vp, u32, u32]; L.h264b200EngineSetWindow.restype = None
    L.h264b200DeviceParse.argtypes = [sp]; L.h264b200DeviceParse.restype = u32
    if hasattr(L, "h264b200DebugFetchPredeblock"):
        L.h264b200DebugFetchPredeblock.argtypes = [sp, vp, ctypes.c_size_t]; L.h264b200DebugFetchPredeblock.restype = ctypes.c_long
    if hasattr(L, "h264b200DebugFetchParse"):
        L.h264b200DebugFetchParse.argtypes = [sp, ctypes.c_int, vp, vp, u32, vp]; L.h264b200DebugFetchParse.restype = ctypes.c_int
    L.h264b200DecodeStreams.argtypes = [vp, ctypes.POINTER(StreamDesc), u32, u32, vp, vp, ctypes.POINTER(RunStats)]
    L.h264b200DecodeStreams.restype = ctypes.c_int
    L.h264b200SplitGops.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                    ctypes.POINTER(ctypes.c_size_t), u32]
    L.h264b200SplitGops.restype = ctypes.c_int
    L.h264b200Mp4ToAnnexB.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    L.h264b200Mp4ToAnnexB.restype = ctypes.c_long
    _lib = L
    return L


def probe():
    """(rc, message) of h264b200Probe: rc 0 means an sm_100 CUDA device is usable."""
    buf = ctypes.create_string_buffer(256)
    rc = lib().h264b200Probe(buf, 256)
    return rc, buf.value.decode()


def require_gpu():
    rc, msg = probe()
    if rc != 0:
        raise RuntimeError("libh264b200.so has no CPU reconstruction path and CUDA is unusable: " + msg)
    return msg


def frame_md5(ptr, nbytes):
    return hashlib.md5(ctypes.string_at(ptr, nbytes)).hexdigest()


def decode_annexb(data, keep_frames=False, no_reordering=0, api="swdec"):
    """Decode one Annex-B stream through the reference-shaped C entry points.

    api="swdec": H264SwDecInit/Decode/NextPicture/Release, the loop of DecTestBench.c:213-400.
    api="bsd":   h264bsdInit/Decode/NextOutputPicture/Shutdown directly (h264bsd_decoder.h:60-66).
    Returns (list of md5 hex digests or raw frames, info dict)."""
    L = lib()
    require_gpu()
    buf = ctypes.create_string_buffer(bytes(data), len(data) + 16)
    base = ctypes.addressof(buf)
    out = []
    info = {"err_mbs": 0, "width": 0, "height": 0, "pic_ids": []}

    def take(ptr, w, h, pic_id, err):
        n = w * h * 3 // 2
        out.append(ctypes.string_at(ptr, n) if keep_frames else frame_md5(ptr, n))
        info["err_mbs"] += err
        info["pic_ids"].append(pic_id)

    if api == "swdec":
        inst = ctypes.c_void_p()
        if L.H264SwDecInit(ctypes.byref(inst), no_reordering) != H264SWDEC_OK:
            raise RuntimeError("H264SwDecInit failed")
        try:
            inp, outp, pic, inf = SwDecInput(), SwDecOutput(), SwDecPicture(), SwDecInfo()
            pos, n, pic_id = 0, len(data), 0
            while pos < n:
                inp.pStream = base + pos; inp.dataLen = n - pos; inp.picId = pic_id
                ret = L.H264SwDecDecode(inst, ctypes.byref(inp), ctypes.byref(outp))
                if ret < 0:
                    if ret == -4:
                        raise RuntimeError("H264SwDecDecode: H264SWDEC_MEMFAIL (CUDA engine unavailable?)")
                    break
                pos = outp.pStrmCurrPos - base
                if ret == H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY:
                    L.H264SwDecGetInfo(inst, ctypes.byref(inf))
                    info["width"], info["height"] = inf.picWidth, inf.picHeight
                if ret in (H264SWDEC_PIC_RDY, H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY):
                    pic_id += 1
                    while L.H264SwDecNextPicture(inst, ctypes.byref(pic), 0) == H264SWDEC_PIC_RDY:
                        take(pic.pOutputPicture, info["width"], info["height"], pic.picId, pic.nbrOfErrMBs)
            while L.H264SwDecNextPicture(inst, ctypes.byref(pic), 1) == H264SWDEC_PIC_RDY:
                take(pic.pOutputPicture, info["width"], info["height"], pic.picId, pic.nbrOfErrMBs)
        finally:
            L.H264SwDecRelease(inst)
    else:
        st = Storage()
        if L.h264bsdInit(ctypes.byref(st), no_reordering) != 0:
            raise RuntimeError("h264bsdInit failed")
        try:
            pos, n, pic_id = 0, len(data), 0
            nread, pid, idr, err = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()

            def drain():
                while True:
                    p = L.h264bsdNextOutputPicture(ctypes.byref(st), ctypes.byref(pid), ctypes.byref(idr), ctypes.byref(err))
                    if not p:
                        break
                    take(p, info["width"], info["height"], pid.value, err.value)
            while pos < n:
                rc = L.h264bsdDecode(ctypes.byref(st), base + pos, n - pos, pic_id, ctypes.byref(nread))
                pos += nread.value
                if rc == H264BSD_MEMALLOC_ERROR:
                    raise RuntimeError("h264bsdDecode: H264BSD_MEMALLOC_ERROR (CUDA engine unavailable?)")
                if rc == H264BSD_HDRS_RDY:
                    info["width"] = 16 * L.h264bsdPicWidth(ctypes.byref(st)); info["height"] = 16 * L.h264bsdPicHeight(ctypes.byref(st))
                    drain()
                elif rc == H264BSD_PIC_RDY:
                    pic_id += 1
                    drain()
                elif nread.value == 0:
                    break
            L.h264bsdFlushBuffer(ctypes.byref(st))
            drain()
        finally:
            L.h264bsdShutdown(ctypes.byref(st))
    return out, info


def decode_on_engine(eng, data, keep_frames=False, fetch_parse=False, no_reordering=0, predeblock=None):
    """Decode one Annex-B stream with the h264bsd* loop on an instance attached to `eng` (h264b200InitOnEngine).  On an
    engine that is not batched every finished picture is launched at once, so this is the synchronous API on an engine
    of the caller's choice — in particular a device-parse one (ENGINE_DEVICE_PARSE): slice data parsed by kernel Kp.
    fetch_parse: also return, per picture in decoding order, what Kp produced (h264b200DebugFetchParse):
    (records bytes, coefficient slot bytes, 12 result words).  predeblock: a list that receives, per picture in decoding
    order, the MD5 of the picture before deblocking (engine flag ENGINE_TAP_PREDEBLOCK).
    Returns (md5s or frames, info[, parses])."""
    L = lib()
    buf = ctypes.create_string_buffer(bytes(data), len(data) + 16)
    base = ctypes.addressof(buf)
    out, parses = [], []
    info = {"err_mbs": 0, "width": 0, "height": 0, "pic_ids": [], "device_parse": 0}
    st = Storage()
    if L.h264b200InitOnEngine(ctypes.byref(st), no_reordering, eng.h) != 0:
        raise RuntimeError("h264b200InitOnEngine failed")
    try:
        pos, n, pic_id = 0, len(data), 0
        nread, pid, idr, err = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()

        def drain():
            while True:
                p = L.h264bsdNextOutputPicture(ctypes.byref(st), ctypes.byref(pid), ctypes.byref(idr), ctypes.byref(err))
                if not p:
                    break
                nb = info["width"] * info["height"] * 3 // 2
                out.append(ctypes.string_at(p, nb) if keep_frames else frame_md5(p, nb))
                info["err_mbs"] += err.value
                info["pic_ids"].append(pid.value)

        def fetch_pre():
            nb = info["width"] * info["height"] * 3 // 2
            buf = ctypes.create_string_buffer(nb)
            rc = L.h264b200DebugFetchPredeblock(ctypes.byref(st), buf, nb)
            if rc != nb:
                raise RuntimeError("h264b200DebugFetchPredeblock failed (%d)" % rc)
            predeblock.append(hashlib.md5(buf.raw).hexdigest())

        def fetch():
            n_mbs = info["width"] * info["height"] // 256
            mbs = ctypes.create_string_buffer(n_mbs * 128)
            cap = n_mbs * 28 + 16
            coef = ctypes.create_string_buffer(cap * 32)
            res = (ctypes.c_uint32 * 12)()
            rc = L.h264b200DebugFetchParse(ctypes.byref(st), 0, mbs, coef, cap, res)
            if rc != 0:
                raise RuntimeError("h264b200DebugFetchParse failed (%d)" % rc)
            parses.append((mbs.raw, coef.raw[:res[0] * 32], list(res)))
        while pos < n:
            rc = L.h264bsdDecode(ctypes.byref(st), base + pos, n - pos, pic_id, ctypes.byref(nread))
            pos += nread.value
            if rc == H264BSD_MEMALLOC_ERROR:
                raise RuntimeError("h264bsdDecode: H264BSD_MEMALLOC_ERROR (CUDA engine unavailable?)")
            if rc == H264BSD_HDRS_RDY:
                info["width"] = 16 * L.h264bsdPicWidth(ctypes.byref(st)); info["height"] = 16 * L.h264bsdPicHeight(ctypes.byref(st))
                drain()
            elif rc == H264BSD_PIC_RDY:
                pic_id += 1
                info["device_parse"] = L.h264b200DeviceParse(ctypes.byref(st))
                if fetch_parse:
                    fetch()
                if predeblock is not None:
                    fetch_pre()
                drain()
            elif nread.value == 0:
                break
        before = len(info["pic_ids"])
        L.h264bsdFlushBuffer(ctypes.byref(st))
        if predeblock is not None and L.h264b200DeviceParse(ctypes.byref(st)) and len(predeblock) < pic_id + 1:
            fetch_pre()                 # device-parse: the flush ended (and launched) the last picture
        if fetch_parse and L.h264b200DeviceParse(ctypes.byref(st)) and len(parses) < pic_id + 1:
            try:
                fetch()                 # the picture the flush ended (device-parse: the last one of the stream)
            except RuntimeError:
                pass
        drain()
        del before
    finally:
        L.h264bsdShutdown(ctypes.byref(st))
    return (out, info, parses) if fetch_parse else (out, info)


def split_gops(data, max_segs=4096):
    """h264b200SplitGops: list of self-contained IDR-bounded segments (bytes)."""
    L = lib()
    cap = len(data) * 2 + 4096
    src = ctypes.create_string_buffer(bytes(data), len(data))
    dst = ctypes.create_string_buffer(cap)
    off = (ctypes.c_size_t * max_segs)()
    ln = (ctypes.c_size_t * max_segs)()
    n = L.h264b200SplitGops(ctypes.addressof(src), len(data), ctypes.addressof(dst), cap, off, ln, max_segs)
    if n < 0:
        raise RuntimeError("h264b200SplitGops failed")
    return [dst.raw[off[i]:off[i] + ln[i]] for i in range(n)]


def mp4_to_annexb(mp4):
    """h264b200Mp4ToAnnexB: (annexb bytes, number of samples)."""
    L = lib()
    src = ctypes.create_string_buffer(bytes(mp4), len(mp4))
    cap = len(mp4) + 65536
    dst = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t()
    rc = L.h264b200Mp4ToAnnexB(ctypes.addressof(src), len(mp4), ctypes.addressof(dst), cap, ctypes.byref(n))
    if rc < 0:
        raise ValueError("h264b200Mp4ToAnnexB failed (%d)" % rc)
    return dst.raw[:n.value], rc


class Engine:
    """One batch engine on one CUDA device (h264b200EngineCreateEx)."""

    def __init__(self, device=-1, flags=ENGINE_BATCHED):
        require_gpu()
        self.L = lib()
        self.h = self.L.h264b200EngineCreateEx(device, flags)
        if not self.h:
            raise RuntimeError("h264b200EngineCreateEx failed (see stderr)")

    def close(self):
        if self.h:
            self.L.h264b200EngineDestroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_flags(self, flags):
        self.L.h264b200EngineSetFlags(self.h, flags)

    def decode_streams(self, streams, threads=0, on_picture=None):
        """Decode independent Annex-B streams, one picture of each per batched launch.
        on_picture(stream, index, ptr, width, height, pic_id, err_mbs) is called from worker threads."""
        n = len(streams)
        # the C side never writes to the caller's streams (it makes its own copies, in the worker
        # threads), so the bytes objects are passed by address: no copy here
        keep = [s if isinstance(s, bytes) else bytes(s) for s in streams]
        descs = (StreamDesc * n)()
        for i, b in enumerate(keep):
            descs[i].data = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p); descs[i].len = len(b)
        cb = None
        if on_picture is not None:
            def _cb(user, stream, index, ptr, w, h, pic_id, err):
                on_picture(stream, index, ptr, w, h, pic_id, err)
            cb = PICTURE_CB(_cb)
        rs = RunStats()
        rc = self.L.h264b200DecodeStreams(self.h, descs, n, threads, ctypes.cast(cb, ctypes.c_void_p) if cb else None, None, ctypes.byref(rs))
        if rc != 0:
            raise RuntimeError("h264b200DecodeStreams failed (rc=%d, failed streams=%d)" % (rc, rs.failed_streams))
        return rs

    def decode_streams_md5(self, streams, threads=0):
        """Per-stream list of per-picture MD5 digests (Y|Cb|Cr of the MB-aligned frame)."""
        res = [dict() for _ in streams]

        def on_picture(stream, index, ptr, w, h, pic_id, err):
            res[stream][index] = frame_md5(ptr, w * h * 3 // 2)
        rs = self.decode_streams(streams, threads, on_picture)
        return [[d[i] for i in sorted(d)] for d in res], rs

    def replay(self, reps=1, time_kernels=True):
        return self.L.h264b200EngineReplay(self.h, reps, 1 if time_kernels else 0)

    def replay_ms(self):
        return self.L.h264b200EngineReplayMs(self.h)

    def sync(self):
        self.L.h264b200EngineSync(self.h)

    def kernel_times(self, reset=False):
        kt = KernelTimes()
        self.L.h264b200EngineKernelTimes(self.h, ctypes.byref(kt), 1 if reset else 0)
        names = ("k1_transform", "k2_inter", "k3_intra", "k4_deblock", "kp_parse")
        return {names[k]: {"ms": kt.ms[k], "bytes": kt.bytes[k], "launches": kt.launches[k]} for k in range(5)}

    def check_resident(self):
        return self.L.h264b200EngineCheckResident(self.h)

    def drop_retained(self):
        self.L.h264b200EngineDropRetained(self.h)

    def stats(self):
        s = Stats()
        self.L.h264b200EngineStats(self.h, ctypes.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def error_flags(self):
        return self.L.h264b200EngineErrorFlags(self.h)
