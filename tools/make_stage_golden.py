"""Stage goldens made with the UNMODIFIED reference (oracle/_ref/libh264ref.so, built by oracle/Makefile; its only addition
is oracle/ref_tap.c, an observation hook on h264bsdFilterPicture):

  tests/golden/predeblock.json   per picture, in decoding order, the MD5 of the picture BEFORE in-loop deblocking — the output
                                 of K1..K3 alone (h264bsd_decoder.c:489-491 is where the reference filters)
  tests/golden/k1_transform.json for every QP 0..51 the MD5 of h264bsdProcessBlock / h264bsdProcessLumaDc /
                                 h264bsdProcessChromaDc (h264bsd_transform.c:94-398) over seeded random coefficient blocks,
                                 driven the way ProcessResidual does (h264bsd_macroblock_layer.c:1343-1424)

Run here (needs /root/reference):  python tools/make_stage_golden.py"""
import ctypes
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libh264ref.so")


class In(ctypes.Structure):
    _fields_ = [("pStream", ctypes.c_void_p), ("dataLen", ctypes.c_uint32), ("picId", ctypes.c_uint32), ("intraConcealmentMethod", ctypes.c_uint32)]


class Out(ctypes.Structure):
    _fields_ = [("pStrmCurrPos", ctypes.c_void_p)]


def reflib():
    L = ctypes.CDLL(REFLIB)
    L.H264SwDecInit.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32]
    L.H264SwDecDecode.argtypes = [ctypes.c_void_p, ctypes.POINTER(In), ctypes.POINTER(Out)]
    L.H264SwDecRelease.argtypes = [ctypes.c_void_p]
    L.reftap_set_predeblock_buffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.reftap_predeblock_len.restype = ctypes.c_size_t
    L.reftap_pictures_filtered.restype = ctypes.c_uint
    L.h264bsdProcessBlock.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]; L.h264bsdProcessBlock.restype = ctypes.c_uint32
    L.h264bsdProcessLumaDc.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    L.h264bsdProcessChromaDc.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    return L


def reference_predeblock_md5(L, data, n_mbs):
    """MD5 of every picture before deblocking, in decoding order (pictures whose filter is disabled everywhere are
    still handed to h264bsdFilterPicture, so every picture is seen)."""
    buf = ctypes.create_string_buffer(bytes(data), len(data) + 16)
    pre = ctypes.create_string_buffer(n_mbs * 384)
    inst = ctypes.c_void_p()
    assert L.H264SwDecInit(ctypes.byref(inst), 0) == 0
    L.reftap_set_predeblock_buffer(pre, n_mbs * 384)
    out_md5, seen = [], L.reftap_pictures_filtered()
    i, o = In(), Out()
    pos, n, base = 0, len(data), ctypes.addressof(buf)
    try:
        while pos < n:
            i.pStream = base + pos; i.dataLen = n - pos; i.picId = 0
            ret = L.H264SwDecDecode(inst, ctypes.byref(i), ctypes.byref(o))
            if ret < 0:
                break
            pos = o.pStrmCurrPos - base
            now = L.reftap_pictures_filtered()
            if now != seen:
                assert now == seen + 1
                seen = now
                out_md5.append(hashlib.md5(pre.raw[:L.reftap_predeblock_len()]).hexdigest())
    finally:
        L.reftap_set_predeblock_buffer(None, 0)
        L.H264SwDecRelease(inst)
    return out_md5


# ---- K1: the reference's residual processing of one macroblock, from the three transform functions
ZIGZAG = [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]        # scan position -> raster index
DC_INDEX = [0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15]      # h264bsd_macroblock_layer.c:78-79 dcCoeffIndex
QPC = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33, 34, 34, 35,
       35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]


def k1_cases(qp, seed):
    """Seeded coefficient sets (SCAN order, like the bitstream delivers them) for one Intra16x16-style macroblock."""
    rng = random.Random(seed * 1000 + qp)
    budget = max(1, min(60, int(700 / 2 ** (qp / 6.0))))           # sum of |levels| a block can carry before its residual leaves [-512, 511]
    if seed >= 4:
        budget *= 40                                               # ... and beyond: the range check itself (h264bsd_transform.c:181-185)

    def blk(n, first=0):
        v = [0] * 16
        n = min(n, budget, 16 - first)
        amp = max(1, budget // (2 * max(n, 1)))
        for k in rng.sample(range(first, 16), n):
            v[k] = rng.choice([-1, 1]) * rng.randrange(1, amp + 1)
        return v
    luma_dc = blk(rng.randrange(0, 5))
    luma = [blk(rng.choice([0, 0, 1, 2, 5, 15]), 1) for _ in range(16)]
    cdc = [[rng.choice([-1, 0, 1]) * rng.randrange(1, max(1, budget // 6) + 1) for _ in range(4)] for _ in range(2)]
    cac = [blk(rng.choice([0, 1, 3, 15]), 1) for _ in range(8)]
    plain = [blk(rng.choice([0, 1, 2, 4, 16])) for _ in range(16)]  # luma blocks of a macroblock that is not Intra16x16
    return luma_dc, luma, cdc, cac, plain


def reference_k1(L, qp, chroma_off, case):
    """Residual of every block (raster order inside a block, 16 values; None where the reference leaves the block
    empty) as ProcessResidual computes it: I16x16 luma, both chroma planes, and 16 plain luma blocks."""
    luma_dc, luma, cdc, cac, plain = case
    I32x16 = ctypes.c_int32 * 16

    def process_block(scan_vals, q, skip):
        cmap = 0
        for k, v in enumerate(scan_vals):
            if v:
                cmap |= 1 << k
        d = I32x16(*scan_vals)
        rc = L.h264bsdProcessBlock(d, q, skip, cmap)
        return list(d), rc
    out = {"i16": [], "chroma": [], "plain": [], "range_error": False}
    dc = I32x16(*luma_dc)                                         # levels arrive in scan order; the reference un-zig-zags inside
    L.h264bsdProcessLumaDc(dc, qp)
    for b in range(16):
        vals = list(luma[b]); vals[0] = dc[DC_INDEX[b]]
        if any(vals):
            r, rc = process_block(vals, qp, 1)
            out["range_error"] |= bool(rc); out["i16"].append(r)
        else:
            out["i16"].append(None)
    qpc = QPC[max(0, min(51, qp + chroma_off))]
    c = I32x16(*(cdc[0] + cdc[1] + [0] * 8))
    L.h264bsdProcessChromaDc(c, qpc)
    for b in range(8):
        vals = list(cac[b]); vals[0] = c[b]
        if any(vals):
            r, rc = process_block(vals, qpc, 1)
            out["range_error"] |= bool(rc); out["chroma"].append(r)
        else:
            out["chroma"].append(None)
    for b in range(16):
        if any(plain[b]):
            r, rc = process_block(list(plain[b]), qp, 0)
            out["range_error"] |= bool(rc); out["plain"].append(r)
        else:
            out["plain"].append(None)
    return canonical(out)


def canonical(out):
    """what is hashed: empty blocks as zeros (an empty block and an all-zero residual give the same samples); when the
    reference reports a residual outside [-512, 511] (h264bsd_transform.c:181-185) only that fact"""
    if out["range_error"]:
        return {"range_error": True}
    z = [0] * 16
    return {"range_error": False, "i16": [b or z for b in out["i16"]], "chroma": [b or z for b in out["chroma"]], "plain": [b or z for b in out["plain"]]}


def main():
    import cases
    import __graft_entry__
    __graft_entry__.build()
    L = reflib()
    pre = {}
    for case in cases.SMALL + cases.FULL[:2]:
        data = cases.make_stream(case)
        md5s = reference_predeblock_md5(L, data, case[1] * case[2])
        assert len(md5s) == case[3], (case[0], len(md5s))
        pre[case[0]] = {"stream_md5": hashlib.md5(data).hexdigest(), "predeblock_md5": md5s}
        print(case[0], len(md5s), "pictures", file=sys.stderr)
    json.dump(pre, open(os.path.join(ROOT, "tests", "golden", "predeblock.json"), "w"), indent=0, sort_keys=True)
    k1 = {}
    for qp in range(52):
        h = hashlib.md5()
        for seed in range(5):
            for off in (0, -7, 5):
                h.update(json.dumps(reference_k1(L, qp, off, k1_cases(qp, seed)), sort_keys=True).encode())
        k1[str(qp)] = h.hexdigest()
    json.dump({"seeds": 5, "chroma_offsets": [0, -7, 5], "md5_per_qp": k1}, open(os.path.join(ROOT, "tests", "golden", "k1_transform.json"), "w"), indent=0, sort_keys=True)
    print("wrote predeblock.json, k1_transform.json")


if __name__ == "__main__":
    main()
