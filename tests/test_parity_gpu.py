"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI of
libh264b200.so (include/*.h); the checker is the oracle (CPU restatement run live on the same
bytes) and the committed golden MD5s of the unmodified reference.  Bar: bit-exact (frame MD5)."""
import ctypes
import hashlib

import pytest

import cases
import util
from broadway_b200 import bitstream, capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    capi.require_gpu()


@pytest.mark.parametrize("case", cases.SMALL, ids=[c[0] for c in cases.SMALL])
def test_swdec_api_matches_golden_and_oracle(case, golden):
    data = cases.make_stream(case)
    got, info = capi.decode_annexb(data, api="swdec")
    want, _ = util.oracle_md5(data)
    assert info["err_mbs"] == 0
    assert got == want
    assert got == golden[case[0]]["frame_md5"]


@pytest.mark.parametrize("case", cases.SMALL[:6], ids=[c[0] for c in cases.SMALL[:6]])
def test_h264bsd_entry_points_match_golden(case, golden):
    got, info = capi.decode_annexb(cases.make_stream(case), api="bsd")
    assert got == golden[case[0]]["frame_md5"]
    assert info["pic_ids"] == list(range(len(got)))


@pytest.mark.parametrize("case", cases.FULL, ids=[c[0] for c in cases.FULL])
def test_full_size_matches_reference_golden(case, golden):
    got, info = capi.decode_annexb(cases.make_stream(case))
    assert (info["width"], info["height"]) == (16 * case[1], 16 * case[2])
    assert got == golden[case[0]]["frame_md5"]


def test_batched_streams_match_golden(golden):
    """All small cases at once through one engine: pictures of different sizes and types share launches."""
    streams = [cases.make_stream(c) for c in cases.SMALL]
    with capi.Engine() as eng:
        md5s, rs = eng.decode_streams_md5(streams, threads=4)
        assert rs.failed_streams == 0 and rs.err_mbs == 0
        for c, m in zip(cases.SMALL, md5s):
            assert m == golden[c[0]]["frame_md5"], c[0]
        st = eng.stats()
        assert st["pictures"] == sum(c[3] for c in cases.SMALL)
        # the runner launches the streams as two groups per round: at most two batches per picture index
        assert st["batches"] <= 2 * (max(c[3] for c in cases.SMALL) + 2)
        assert eng.error_flags() == 0


def test_thread_count_and_batch_shape_do_not_change_results(golden):
    streams = [cases.make_stream(c) for c in cases.SMALL[:10]]
    ref = None
    for threads in (1, 3, 10):
        with capi.Engine() as eng:
            md5s, _ = eng.decode_streams_md5(streams, threads=threads)
        if ref is None:
            ref = md5s
        assert md5s == ref


def test_gop_segments_in_parallel_equal_sequential_decode():
    """Size-independent property at 1080p: IDR-bounded segments decoded as independent instances in
    one batch give exactly the frames of the sequential decode."""
    data = bitstream.synth(120, 68, 12, seed=4242, idr_period=3)
    seq, info = capi.decode_annexb(data)
    assert len(seq) == 12
    segs = capi.split_gops(data)
    assert len(segs) == 4
    with capi.Engine() as eng:
        md5s, rs = eng.decode_streams_md5(segs, threads=4)
    assert [m for s in md5s for m in s] == seq
    # and against the oracle on the first segment (the CPU restatement needs ~0.1 s per 1080p frame)
    want, _ = util.oracle_md5(segs[0])
    assert md5s[0] == want


def test_resident_replay_reproduces_the_pictures():
    """bench.py's kernel-only leg replays retained batches from HBM: it must rebuild the same frames."""
    streams = [bitstream.synth(20, 12, 5, seed=100 + i, p_intra_permille=100) for i in range(6)]
    with capi.Engine(flags=capi.ENGINE_BATCHED | capi.ENGINE_RETAIN) as eng:
        md5s, _ = eng.decode_streams_md5(streams, threads=2)
        for s, m in zip(streams, md5s):
            assert m == util.oracle_md5(s)[0]
        assert eng.check_resident() == 0
        n = eng.replay(reps=3, time_kernels=True)
        assert n == 3 * 30
        assert eng.check_resident() == 0
        kt = eng.kernel_times()
        assert kt["k2_inter"]["launches"] == 3 * 4 and kt["k4_deblock"]["ms"] > 0


def test_idempotent_redecode_and_instance_reuse():
    data = cases.make_stream(cases.SMALL[0])
    a, _ = capi.decode_annexb(data)
    b, _ = capi.decode_annexb(data)
    assert a == b


def test_output_is_planar_i420_mb_aligned():
    """Decoder.c:113-147 contract: width*height*3/2 bytes, Y then Cb then Cr; an I_PCM picture makes it checkable."""
    data = bitstream.synth(4, 3, 1, seed=5, first_idr_ipcm=1, deblock_idc=1)
    frames, info = capi.decode_annexb(data, keep_frames=True)
    assert (info["width"], info["height"]) == (64, 48) and len(frames[0]) == 64 * 48 * 3 // 2
    ref, _ = util.oracle_md5(data)
    assert hashlib.md5(frames[0]).hexdigest() == ref[0]


def test_bare_nal_input_like_the_mp4_player():
    """Player/mp4.js feeds one bare NAL (no start code) per decode() call (h264bsd_byte_stream.c:101-103)."""
    data = cases.make_stream(cases.SMALL[0])
    nals = [n for n in data.split(b"\x00\x00\x00\x01") if n]
    L = capi.lib()
    st = capi.Storage()
    assert L.h264bsdInit(ctypes.byref(st), 0) == 0
    out, nread = [], ctypes.c_uint32()
    pid, idr, err = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    try:
        for i, nal in enumerate(nals):
            buf = ctypes.create_string_buffer(nal, len(nal) + 8)
            while True:
                rc = L.h264bsdDecode(ctypes.byref(st), ctypes.addressof(buf), len(nal), i, ctypes.byref(nread))
                if rc == capi.H264BSD_HDRS_RDY:
                    w, h = 16 * L.h264bsdPicWidth(ctypes.byref(st)), 16 * L.h264bsdPicHeight(ctypes.byref(st))
                    continue           # same buffer again (readBytes == 0)
                break
            if rc == capi.H264BSD_PIC_RDY:
                while True:
                    p = L.h264bsdNextOutputPicture(ctypes.byref(st), ctypes.byref(pid), ctypes.byref(idr), ctypes.byref(err))
                    if not p:
                        break
                    out.append(capi.frame_md5(p, w * h * 3 // 2))
    finally:
        L.h264bsdShutdown(ctypes.byref(st))
    assert out == util.oracle_md5(data)[0]


def test_broadway_shim_play_stream(golden):
    """The Decoder.c surface: fill the stream buffer, broadwayPlayStream, pictures arrive by callback."""
    L = capi.lib()
    data = cases.make_stream(cases.SMALL[0])
    got, hdr = [], []
    HCB = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
    PCB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32)
    on_h = HCB(lambda user: hdr.append(1))
    on_p = PCB(lambda user, p, w, h: got.append(capi.frame_md5(p, w * h * 3 // 2)))
    L.broadwaySetCallbacks.argtypes = [HCB, PCB, ctypes.c_void_p]
    L.broadwayCreateStream.restype = ctypes.c_void_p; L.broadwayCreateStream.argtypes = [ctypes.c_uint32]
    L.broadwayPlayStream.argtypes = [ctypes.c_uint32]
    L.broadwaySetCallbacks(on_h, on_p, None)
    assert L.broadwayInit() == 0
    buf = L.broadwayCreateStream(len(data))
    ctypes.memmove(buf, data, len(data))
    L.broadwayPlayStream(len(data))
    L.broadwayExit()
    L.broadwaySetCallbacks(HCB(), PCB(), None)
    assert hdr == [1]
    assert got == golden[cases.SMALL[0][0]]["frame_md5"]
    assert L.broadwayGetMajorVersion() == 2


def test_mp4_clip_through_the_decoder(golden):
    """Config 1/2 shape (Player/*.mp4 are absent from the mount): MP4 -> h264b200Mp4ToAnnexB -> decode."""
    import mp4mux
    case = next(c for c in cases.SMALL if c[0] == "p_intra_mix")
    mp4, n = mp4mux.mux(cases.make_stream(case), 16 * case[1], 16 * case[2])
    annexb, n2 = capi.mp4_to_annexb(mp4)
    assert n == n2 == case[3]
    got, _ = capi.decode_annexb(annexb)
    assert got == golden["p_intra_mix"]["frame_md5"]


def _rgba_reference(frame, W, H, cl, ct, cw, ch):
    """numpy restatement of templates/DecoderPost.js:514-560 on the cropped rectangle (test oracle)."""
    import numpy as np
    f = np.frombuffer(frame, dtype=np.uint8)
    Y = f[:W * H].reshape(H, W).astype(np.int32)
    U = f[W * H:W * H + W * H // 4].reshape(H // 2, W // 2).astype(np.int32)
    V = f[W * H + W * H // 4:].reshape(H // 2, W // 2).astype(np.int32)
    y = Y[ct:ct + ch, cl:cl + cw]
    ys, xs = np.mgrid[ct:ct + ch, cl:cl + cw]
    u, v = U[ys >> 1, xs >> 1], V[ys >> 1, xs >> 1]
    a0 = 1192 * (y - 16)
    r = np.clip((a0 + 1634 * (v - 128)) >> 10, 0, 255)
    g = np.clip((a0 - 832 * (v - 128) - 400 * (u - 128)) >> 10, 0, 255)
    b = np.clip((a0 + 2066 * (u - 128)) >> 10, 0, 255)
    out = np.stack([r, g, b, np.full_like(r, 255)], axis=-1).astype(np.uint8)
    return out.tobytes()


@pytest.mark.parametrize("crop", [0, 1], ids=["uncropped", "cropped"])
def test_rgba_cropped_output_matches_wrapper_formula(crop):
    """K5: crop + I420->RGBA on the device == the wrapper's converter applied to the I420 frames."""
    w_mbs, h_mbs, n = 7, 5, 3
    data = bitstream.synth(w_mbs, h_mbs, n, seed=77, crop=crop, p_intra_permille=100)
    i420, info = capi.decode_annexb(data, keep_frames=True)
    L = capi.lib()
    L.h264b200SetOutputFormat.argtypes = [ctypes.POINTER(capi.Storage), ctypes.c_uint32]; L.h264b200SetOutputFormat.restype = ctypes.c_uint32
    st = capi.Storage()
    assert L.h264bsdInit(ctypes.byref(st), 0) == 0
    assert L.h264b200SetOutputFormat(ctypes.byref(st), 1) == 0
    buf = ctypes.create_string_buffer(data, len(data) + 16)
    pos, nread, got = 0, ctypes.c_uint32(), []
    pid, idr, err = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    cf, cl, cw, ct, ch = (ctypes.c_uint32() for _ in range(5))
    W, H = 16 * w_mbs, 16 * h_mbs
    try:
        while pos < len(data):
            rc = L.h264bsdDecode(ctypes.byref(st), ctypes.addressof(buf) + pos, len(data) - pos, 0, ctypes.byref(nread))
            pos += nread.value
            if rc == capi.H264BSD_HDRS_RDY:
                L.h264bsdCroppingParams(ctypes.byref(st), ctypes.byref(cf), ctypes.byref(cl), ctypes.byref(cw), ctypes.byref(ct), ctypes.byref(ch))
                rect = (cl.value, ct.value, cw.value, ch.value) if cf.value else (0, 0, W, H)
            elif rc == capi.H264BSD_PIC_RDY:
                while True:
                    p = L.h264bsdNextOutputPicture(ctypes.byref(st), ctypes.byref(pid), ctypes.byref(idr), ctypes.byref(err))
                    if not p:
                        break
                    got.append(ctypes.string_at(p, rect[2] * rect[3] * 4))
            elif nread.value == 0:
                break
    finally:
        L.h264bsdShutdown(ctypes.byref(st))
    assert cf.value == crop and len(got) == n
    if crop:
        assert rect[3] == H - 8
    for a, f in zip(got, i420):
        assert a == _rgba_reference(f, W, H, *rect)


def test_arbitrary_slice_order_on_gpu(golden):
    """ASO: the wavefront kernels take availability from the records, not from arrival order."""
    for name in ("multi_slice", "deblock_idc2"):
        case = next(c for c in cases.SMALL if c[0] == name)
        got, info = capi.decode_annexb(cases.reverse_slice_order(cases.make_stream(case)))
        assert info["err_mbs"] == 0 and got == golden[name]["frame_md5"]


@pytest.mark.parametrize("lc", cases.LOSS, ids=[c[0] for c in cases.LOSS])
def test_concealment_of_lost_slices_on_gpu(lc):
    """Lost slice NAL units: P macroblocks are copied from the reference picture by K2, I macroblocks are
    interpolated by k3c_conceal in the reference's order, both deblocked as intra with QP 40 — frames and the
    concealed-macroblock count equal the reference's (tests/golden/loss.json)."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.json")))[lc[0]]
    got, info = capi.decode_annexb(cases.make_loss_stream(lc))
    assert got == g["frame_md5"]
    assert info["err_mbs"] == g["err_mbs"]


def test_concealment_in_a_batch():
    """Lossy and clean streams share launches; the conceal kernel only touches the pictures that need it."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.json")))
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")))
    streams = [cases.make_loss_stream(lc) for lc in cases.LOSS] + [cases.make_stream(c) for c in cases.SMALL[:4]]
    with capi.Engine() as eng:
        md5s, rs = eng.decode_streams_md5(streams, threads=4)
    for lc, m in zip(cases.LOSS, md5s):
        assert m == g[lc[0]]["frame_md5"], lc[0]
    for c, m in zip(cases.SMALL[:4], md5s[len(cases.LOSS):]):
        assert m == gold[c[0]]["frame_md5"], c[0]
    assert rs.err_mbs == sum(g[lc[0]]["err_mbs"] for lc in cases.LOSS)


def test_corrupted_streams_never_fault_and_equal_the_oracle():
    """Bit errors inside slices: the CUDA path must not fault or hang on whatever records survive, and (same host
    parser, same records) must produce the oracle's pictures and concealed-macroblock counts."""
    import random
    rng = random.Random(55)
    bases = [cases.make_stream(c) for c in cases.SMALL[:14]]
    for _ in range(24):
        data = bytearray(rng.choice(bases))
        for _ in range(rng.randrange(1, 4)):
            data[rng.randrange(60, len(data))] ^= 1 << rng.randrange(8)
        got, info = capi.decode_annexb(bytes(data))
        want, s = util.oracle_md5(bytes(data))
        assert got == want and info["err_mbs"] == s["err_mbs"]


def test_cuda_luma_equals_ffmpeg_golden():
    """Third-decoder check on the GPU path: the luma plane the CUDA engine reconstructs equals FFmpeg's for every
    case FFmpeg can decode (tests/golden/ffmpeg_luma.json, tools/make_ffmpeg_golden.py; both parse modes)."""
    import json
    import os
    import test_ffmpeg_crosscheck_cpu as ff
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ffmpeg_luma.json")))
    with capi.Engine(flags=capi.ENGINE_DEVICE_PARSE) as dev, capi.Engine(flags=0) as host:
        for case in ff.CASES:
            if case[1] * case[2] > 8160:
                continue                                   # 4K: covered by the MD5 goldens
            g = fx[case[0]]
            data = cases.make_stream(case)
            for eng in (dev, host):
                frames, info = capi.decode_on_engine(eng, data, keep_frames=True)
                got = ff.luma_md5s(b"".join(frames), info["width"], info["height"], g["rows"], g["cols"])
                assert got == g["luma_md5"], case[0]


@pytest.mark.parametrize("cap", [1, 4, 16])
def test_wavefront_handover_stress(cap, monkeypatch):
    """the K3 / K4 row hand-over (st.release -> acquire poll, k_common.cuh) under stress: 200
    replays of a retained batch of pictures of mixed sizes and types, at 1, 4 and 16 wavefront CTAs per SM (few CTAs:
    long ticket queues, rows of one picture on the same warps; many: every row pair of every picture in flight at
    once).  Every replay must rebuild exactly the frames of the live decode, which were checked against the goldens."""
    monkeypatch.setenv("H264B200_WF_CAP", str(cap))
    sel = [c for c in cases.SMALL if c[0] in ("ippp_default", "intra_only", "p_intra_mix", "multi_slice", "one_row", "one_col",
                                               "odd_size", "wide", "deblock_idc2", "constrained_intra")]
    streams = [cases.make_stream(c) for c in sel] + [bitstream.synth(120, 68, 3, seed=77 + i, p_intra_permille=200) for i in range(3)]
    with capi.Engine(flags=capi.ENGINE_BATCHED | capi.ENGINE_RETAIN | capi.ENGINE_DEVICE_PARSE) as eng:
        eng.decode_streams_md5(streams, threads=4)
        assert eng.check_resident() == 0
        for _ in range(20):
            eng.replay(reps=10, time_kernels=False)
            assert eng.check_resident() == 0
        assert eng.error_flags() == 0
