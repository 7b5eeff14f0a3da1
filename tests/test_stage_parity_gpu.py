"""GPU stage parity (run with -m gpu): the CUDA kernels checked one stage at a time through the C ABI's parity taps.

  K1       coefficient slots after k1_transform (h264b200DebugFetchParse on an engine that reconstructs: K1 works in place)
           == the oracle's residual tap, on streams covering every QP 0..51 — the oracle's transform is pinned to
           h264bsdProcessBlock / LumaDc / ChromaDc for all 52 QPs by tests/test_stage_parity_cpu.py;
  K1..K3   the picture between k3_intra and k4_deblock (engine flag ENGINE_TAP_PREDEBLOCK) == the reference's picture at its
           h264bsdFilterPicture call (tests/golden/predeblock.json, made with oracle/ref_tap.c), every picture of every case;
  K4       follows: the final pictures equal the reference goldens (tests/test_parity_gpu.py) while the inputs of K4 do."""
import ctypes
import hashlib
import json
import os

import pytest

import cases
import util
from broadway_b200 import bitstream, capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = json.load(open(os.path.join(ROOT, "tests", "golden", "predeblock.json")))
PRE_CASES = [c for c in cases.SMALL + cases.FULL[:2] if c[0] in PRE]


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    capi.require_gpu()


@pytest.mark.parametrize("parse", ["host", "device"])
def test_predeblock_picture_equals_reference(parse):
    flags = capi.ENGINE_TAP_PREDEBLOCK | (capi.ENGINE_DEVICE_PARSE if parse == "device" else 0)
    with capi.Engine(flags=flags) as eng:
        for case in PRE_CASES:
            data = cases.make_stream(case)
            assert hashlib.md5(data).hexdigest() == PRE[case[0]]["stream_md5"]
            pre = []
            capi.decode_on_engine(eng, data, predeblock=pre)
            assert pre == PRE[case[0]]["predeblock_md5"], case[0]


def oracle_residual(data):
    """per picture, in decoding order, the coefficient slots after the oracle's K1 (oracle/recon_cpu.h residual tap)"""
    L = util.cpuchk_lib()
    out = []
    RES_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32)
    cb = RES_CB(lambda user, coef, n: out.append(ctypes.string_at(coef, n * 32)))
    tap = util.Tap(None, ctypes.cast(None, util.TAP_RECORDS), ctypes.cast(cb, ctypes.c_void_p), None)
    L.recon_cpu_set_tap(ctypes.byref(tap))
    eng = L.h264b200EngineCreateEx(0, 0)
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
    return out


def test_k1_residual_equals_oracle_for_every_qp():
    n_blocks = 0
    with capi.Engine(flags=capi.ENGINE_DEVICE_PARSE) as eng:
        for qp in range(52):
            lvl = max(1, min(6, 40 >> (qp // 6)))
            data = bitstream.synth(5, 4, 2, seed=700 + qp, qp=qp, qp_jitter=0, coded_blk_permille=600, max_coeffs=6, max_level=lvl,
                                   p_intra_permille=300, chroma_qp_index_offset=(qp % 5) - 2)
            want = oracle_residual(data)
            _, info, parses = capi.decode_on_engine(eng, data, fetch_parse=True)
            assert info["err_mbs"] == 0 and len(parses) == len(want) == 2
            for (mbs, coef, res), w in zip(parses, want):
                assert coef == w, "residual after K1 differs at QP %d" % qp
                n_blocks += len(w) // 32
        assert eng.error_flags() == 0
    assert n_blocks > 5000
