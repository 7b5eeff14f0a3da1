"""Parity cases shared by tools/make_golden.py (which runs the UNMODIFIED reference
over them and commits per-frame MD5s to tests/golden/streams.json) and the tests.

Every case is a deterministic synthetic Baseline stream from the in-repo writer
(include/h264b200_writer.h); `kw` are h264w_params_t overrides.  They cover the
edge cases the reference's own decode loop distinguishes: intra-only / IPPP /
I_PCM, every partition shape, P_Skip runs, intra in P pictures with and without
constrained_intra_pred, several reference frames, several slices per picture
with per-slice deblocking controls (idc 0/1/2, offsets), QP changes, chroma QP
offset, far out-of-picture vectors (edge clamp), POC types 0 and 2, cropping,
the degenerate one-macroblock picture, escape-coded levels, and decoded-picture-buffer
handling (non-reference pictures, frame_num wrap, list reordering, MMCO 1/2/4/6,
long-term pictures), and flexible macroblock ordering (all seven map types).
"""

# name, width_mbs, height_mbs, frames, overrides
SMALL = [
    ("ippp_default",      20, 12, 6, dict()),
    ("intra_only",        20, 12, 3, dict(intra_only=1)),
    ("intra_nodeblock",   20, 12, 2, dict(intra_only=1, deblock_idc=1)),
    ("ipcm_then_p16",     20, 12, 4, dict(first_idr_ipcm=1, part_mix=0, coded_blk_permille=0, p_intra_permille=0, p_skip_permille=0)),
    ("p_noresidual_nodbk", 20, 12, 4, dict(first_idr_ipcm=1, coded_blk_permille=0, p_intra_permille=0, deblock_idc=1)),
    ("p_dense_residual",  20, 12, 4, dict(coded_blk_permille=900, max_coeffs=16, max_level=3)),
    ("p_intra_mix",       20, 12, 6, dict(p_intra_permille=250, ipcm_permille=100)),
    ("constrained_intra", 20, 12, 5, dict(p_intra_permille=300, constrained_intra_pred=1)),
    ("multi_ref",         20, 12, 8, dict(num_ref_frames=4)),
    ("multi_slice",       20, 12, 5, dict(slices_per_pic=4, multi_slice_params=1, qp_jitter=6, p_intra_permille=100)),
    ("deblock_idc2",      20, 12, 4, dict(slices_per_pic=3, deblock_idc=2, p_intra_permille=100)),
    ("deblock_offsets",   20, 12, 4, dict(alpha_c0_offset_div2=3, beta_offset_div2=-2, qp=34)),
    ("chroma_qp_offset",  20, 12, 4, dict(chroma_qp_index_offset=-5, qp_jitter=8)),
    ("low_qp",            20, 12, 3, dict(qp=4, qp_jitter=4, max_level=2)),
    ("high_qp",           20, 12, 3, dict(qp=46, qp_jitter=5, max_level=1, max_coeffs=2)),
    ("far_mv",            20, 12, 5, dict(far_mv_permille=200, mv_range_qpel=300)),
    ("skip_heavy",        20, 12, 5, dict(p_skip_permille=700)),
    ("poc_type0",         20, 12, 6, dict(poc_type=0)),
    ("idr_period",        20, 12, 9, dict(idr_period=3)),
    ("crop",              20, 12, 3, dict(crop=1)),
    ("one_mb",             1,  1, 4, dict()),
    ("one_row",            7,  1, 4, dict(p_intra_permille=200)),
    ("one_col",            1,  6, 4, dict(p_intra_permille=200)),
    ("odd_size",          11,  9, 5, dict(p_intra_permille=150, slices_per_pic=2)),
    ("wide",              45,  3, 3, dict(p_intra_permille=150)),
    ("big_levels",        20, 12, 4, dict(max_level=200, qp=2, qp_jitter=0, max_coeffs=3, coded_blk_permille=500)),
    ("escape_levels",     20, 12, 4, dict(max_level=2000, qp=0, qp_jitter=0, max_coeffs=1, coded_blk_permille=400)),
    ("many_coeffs",       20, 12, 3, dict(max_level=6, max_coeffs=16, coded_blk_permille=700, qp=20)),
    # decoded picture buffer: non-reference pictures, frame_num wrap (24 > 16), list reordering, MMCO 1
    ("dpb_stress",        20, 12, 24, dict(num_ref_frames=3, dpb_stress=1)),
    ("dpb_stress_poc0",   11,  9, 24, dict(num_ref_frames=4, dpb_stress=1, poc_type=0, slices_per_pic=2, p_intra_permille=100)),
    # ... plus long-term reference pictures (MMCO 4 + 6, released by MMCO 2)
    ("dpb_long_term",     11,  9, 40, dict(num_ref_frames=3, dpb_stress=2)),
    ("dpb_long_term_5",   11,  9, 40, dict(num_ref_frames=5, dpb_stress=2, poc_type=0, p_intra_permille=60)),
    # flexible macroblock ordering: the seven slice group map types
    ("fmo_interleaved",   11,  9, 4, dict(fmo_type=1, fmo_groups=3, p_intra_permille=150)),
    ("fmo_dispersed",     11,  9, 4, dict(fmo_type=2, fmo_groups=4, slices_per_pic=2, multi_slice_params=1)),
    ("fmo_foreground",    11,  9, 4, dict(fmo_type=3, fmo_groups=3, p_intra_permille=150)),
    ("fmo_box_out",       11,  9, 5, dict(fmo_type=4, p_intra_permille=100)),
    ("fmo_raster",        11,  9, 5, dict(fmo_type=5, deblock_idc=2)),
    ("fmo_wipe",          11,  9, 5, dict(fmo_type=6, p_intra_permille=100, constrained_intra_pred=1)),
    ("fmo_explicit",      11,  9, 4, dict(fmo_type=7, fmo_groups=8, p_intra_permille=150, num_ref_frames=2)),
]

# BASELINE.json's full-size configurations (few frames: the reference runs at ~20 fps per core)
FULL = [
    ("1080p_ippp",       120, 68, 4, dict()),
    ("1080p_intra",      120, 68, 2, dict(intra_only=1)),
    ("4k_ippp",          240, 135, 3, dict(level_idc=51)),
]

SEED = 20260718


def make_stream(case):
    from broadway_b200 import bitstream
    name, w, h, n, kw = case
    return bitstream.synth(w, h, n, seed=SEED + sum(map(ord, name)), **kw)


def reverse_slice_order(data):
    """Arbitrary slice order (ASO, Baseline): the slices of every picture in reverse order.  Slice membership,
    not arrival order, decides neighbour availability, so the decoded pictures do not change."""
    sc = b"\x00\x00\x00\x01"
    nals = [n for n in data.split(sc) if n]
    out, cur = [], []
    for n in nals:
        if n[0] & 31 in (1, 5):
            if (n[1] & 0x80) and cur:
                out.extend(reversed(cur)); cur = []
            cur.append(n)
        else:
            out.extend(reversed(cur)); cur = []
            out.append(n)
    out.extend(reversed(cur))
    return b"".join(sc + n for n in out)


def drop_slices(data, drops):
    """Remove whole slice NAL units: `drops` is a set of (picture index, slice index within the picture).
    Models packet loss; the decoder must conceal the macroblocks of the missing slices."""
    sc = b"\x00\x00\x00\x01"
    nals = [n for n in data.split(sc) if n]
    out, pic, sl = [], -1, 0
    for n in nals:
        if n[0] & 31 in (1, 5):
            if n[1] & 0x80:            # first_mb_in_slice == 0: a new picture starts
                pic += 1; sl = 0
            keep = (pic, sl) not in drops
            sl += 1
            if not keep:
                continue
        out.append(n)
    return b"".join(sc + n for n in out)


# Packet-loss cases: (name, writer case, set of dropped (picture, slice) NAL units).  The reference conceals the
# macroblocks of the missing slices (h264bsd_conceal.c); goldens also pin the number of concealed macroblocks.
_LOSS_P = ("loss_base_p", 11, 9, 7, dict(slices_per_pic=3, num_ref_frames=2))
_LOSS_I = ("loss_base_i", 11, 9, 5, dict(slices_per_pic=4, intra_only=1))
_LOSS_M = ("loss_base_mixed", 20, 12, 6, dict(slices_per_pic=5, idr_period=3, p_intra_permille=100))
LOSS = [
    ("loss_p_mid_slices",    _LOSS_P, {(2, 1), (4, 1)}),
    ("loss_p_first_slice",   _LOSS_P, {(3, 0)}),
    ("loss_p_two_slices",    _LOSS_P, {(2, 1), (2, 2)}),
    ("loss_whole_picture",   _LOSS_P, {(3, 0), (3, 1), (3, 2)}),
    ("loss_idr_mid_slice",   _LOSS_P, {(0, 1)}),
    ("loss_idr_first_slice", _LOSS_P, {(0, 0)}),
    ("loss_idr_two_slices",  _LOSS_P, {(0, 0), (0, 2)}),
    ("loss_intra_only",      _LOSS_I, {(0, 1), (1, 0), (2, 3), (3, 1), (3, 2)}),
    ("loss_mixed",           _LOSS_M, {(0, 2), (1, 1), (3, 0), (3, 4), (4, 2)}),
]


def make_loss_stream(loss_case):
    name, base, drops = loss_case
    return drop_slices(make_stream(base), drops)
