"""Stage-by-stage pinning of the oracle against the UNMODIFIED reference (VERDICT r1 items 6 / missing 6; SURVEY.md 7.1
steps 3-5), with fixtures made by tools/make_stage_golden.py from oracle/_ref/libh264ref.so:

  K1      oracle/recon_cpu.c recon_cpu_residual_mb == h264bsdProcessBlock / ProcessLumaDc / ProcessChromaDc
          (h264bsd_transform.c:94-398) driven like ProcessResidual (h264bsd_macroblock_layer.c:1343-1424), for every QP 0..51,
          three chroma QP offsets, Intra16x16-style and plain macroblocks (tests/golden/k1_transform.json);
  K1..K3  the oracle's picture BEFORE deblocking == the reference's at its h264bsdFilterPicture call (oracle/ref_tap.c),
          every picture of every case (tests/golden/predeblock.json) — so a deblocking error cannot hide a prediction
          error and vice versa.
The GPU tests (tests/test_stage_parity_gpu.py) compare the CUDA kernels with the same fixtures / the oracle's taps."""
import ctypes
import hashlib
import importlib.util
import json
import os
import struct

import pytest

import cases
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = json.load(open(os.path.join(ROOT, "tests", "golden", "predeblock.json")))
K1 = json.load(open(os.path.join(ROOT, "tests", "golden", "k1_transform.json")))
PRE_CASES = [c for c in cases.SMALL + cases.FULL[:2] if c[0] in PRE]

_spec = importlib.util.spec_from_file_location("make_stage_golden", os.path.join(ROOT, "tools", "make_stage_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

MB_I16, MB_INTER = 2, 0
LUMA_DC, CHROMA_DC = 1 << 24, 1 << 25


def record(mb_class, qp, qpc, mask):
    """one h264b200_mb_t (include/h264b200_records.h): only the fields K1 reads are set"""
    r = bytearray(128)
    r[0], r[1], r[2], r[3] = mb_class, qp, qpc, qp
    struct.pack_into("<II", r, 12, 0, mask)               # coef_offset, resid_mask
    return bytes(r)


def to_slot(scan_vals, first=0):
    """levels in scan order -> the slot layout (raster, un-zig-zagged); `first`: 1 for the 15-coefficient AC blocks"""
    s = [0] * 16
    for k, v in enumerate(scan_vals):
        if k >= first:
            s[gen.ZIGZAG[k]] = v
    return s


def oracle_k1(L, qp, chroma_off, case):
    luma_dc, luma, cdc, cac, plain = case
    qpc = gen.QPC[max(0, min(51, qp + chroma_off))]
    out = {"range_error": False}

    def run(mb_class, mask, slots):
        buf = (ctypes.c_int16 * (16 * max(1, len(slots))))(*[v for s in slots for v in s])
        bad = L.recon_cpu_residual_mb(record(mb_class, qp, qpc, mask), buf)
        out["range_error"] |= bool(bad)
        return [list(buf[16 * i:16 * i + 16]) for i in range(len(slots))]
    # Intra16x16-style: luma DC slot when it has coefficients (then every luma block has a slot), else coded blocks only
    mask, slots, where = 0, [], []
    if any(luma_dc):
        mask |= LUMA_DC | 0xffff
        slots.append(to_slot(luma_dc))
        slots += [to_slot(b, 1) for b in luma]
        where = list(range(1, 17))
    else:
        for b in range(16):
            if any(luma[b][1:]):
                mask |= 1 << b; where.append(len(slots)); slots.append(to_slot(luma[b], 1))
            else:
                where.append(None)
    cwhere = []
    if any(cdc[0]) or any(cdc[1]):
        mask |= CHROMA_DC
        slots.append(cdc[0] + cdc[1] + [0] * 8)
    for b in range(8):
        if any(cac[b][1:]) or any(cdc[b >> 2]):
            mask |= 1 << (16 + b); cwhere.append(len(slots)); slots.append(to_slot(cac[b], 1))
        else:
            cwhere.append(None)
    res = run(MB_I16, mask, slots)
    z = [0] * 16
    out["i16"] = [res[w] if w is not None else z for w in where]
    out["chroma"] = [res[w] if w is not None else z for w in cwhere]
    mask, slots, where = 0, [], []
    for b in range(16):
        if any(plain[b]):
            mask |= 1 << b; where.append(len(slots)); slots.append(to_slot(plain[b]))
        else:
            where.append(None)
    res = run(MB_INTER, mask, slots)
    out["plain"] = [res[w] if w is not None else z for w in where]
    if out["range_error"]:
        return {"range_error": True}
    return {"range_error": False, "i16": out["i16"], "chroma": out["chroma"], "plain": out["plain"]}


@pytest.fixture(scope="module")
def cpuchk():
    L = util.cpuchk_lib()
    L.recon_cpu_residual_mb.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    L.recon_cpu_residual_mb.restype = ctypes.c_int
    return L


@pytest.mark.parametrize("qp", range(52))
def test_oracle_k1_equals_reference_transform(cpuchk, qp):
    h = hashlib.md5()
    for seed in range(K1["seeds"]):
        for off in K1["chroma_offsets"]:
            h.update(json.dumps(oracle_k1(cpuchk, qp, off, gen.k1_cases(qp, seed)), sort_keys=True).encode())
    assert h.hexdigest() == K1["md5_per_qp"][str(qp)]


def test_k1_fixture_is_reproducible_with_the_live_reference():
    if not os.path.exists(gen.REFLIB):
        pytest.skip("oracle/_ref not built here (reference sources absent)")
    L = gen.reflib()
    for qp in (0, 11, 12, 29, 51):
        h = hashlib.md5()
        for seed in range(K1["seeds"]):
            for off in K1["chroma_offsets"]:
                h.update(json.dumps(gen.reference_k1(L, qp, off, gen.k1_cases(qp, seed)), sort_keys=True).encode())
        assert h.hexdigest() == K1["md5_per_qp"][str(qp)]


def oracle_predeblock_md5(data):
    """MD5 of every picture before deblocking, in decoding order, from the oracle's tap (oracle/recon_cpu.h)"""
    from broadway_b200 import capi
    L = util.cpuchk_lib()
    out = []
    PRE_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
    cb = PRE_CB(lambda user, frame, n: out.append(hashlib.md5(ctypes.string_at(frame, n)).hexdigest()))
    tap = util.Tap(None, ctypes.cast(None, util.TAP_RECORDS), None, ctypes.cast(cb, ctypes.c_void_p))
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


@pytest.mark.parametrize("case", PRE_CASES, ids=[c[0] for c in PRE_CASES])
def test_oracle_predeblock_picture_equals_reference(case):
    data = cases.make_stream(case)
    g = PRE[case[0]]
    assert hashlib.md5(data).hexdigest() == g["stream_md5"]
    assert oracle_predeblock_md5(data) == g["predeblock_md5"]
