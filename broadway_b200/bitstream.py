"""ctypes front end of the synthetic H.264 Baseline bitstream writer
(include/h264b200_writer.h, csrc/h264_writer.c).  Test/bench input generator;
not on the decode path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class WriterParams(ctypes.Structure):
    _fields_ = [
        ("width_mbs", ctypes.c_uint32), ("height_mbs", ctypes.c_uint32),
        ("n_frames", ctypes.c_uint32), ("idr_period", ctypes.c_uint32),
        ("intra_only", ctypes.c_uint32), ("seed", ctypes.c_uint64),
        ("qp", ctypes.c_int32), ("qp_jitter", ctypes.c_int32),
        ("coded_blk_permille", ctypes.c_uint32), ("max_coeffs", ctypes.c_uint32),
        ("max_level", ctypes.c_int32), ("num_ref_frames", ctypes.c_uint32),
        ("slices_per_pic", ctypes.c_uint32), ("poc_type", ctypes.c_uint32),
        ("chroma_qp_index_offset", ctypes.c_int32), ("deblock_idc", ctypes.c_uint32),
        ("alpha_c0_offset_div2", ctypes.c_int32), ("beta_offset_div2", ctypes.c_int32),
        ("constrained_intra_pred", ctypes.c_uint32), ("p_intra_permille", ctypes.c_uint32),
        ("p_skip_permille", ctypes.c_uint32), ("ipcm_permille", ctypes.c_uint32),
        ("i16_permille", ctypes.c_uint32), ("mv_range_qpel", ctypes.c_int32),
        ("far_mv_permille", ctypes.c_uint32), ("level_idc", ctypes.c_uint32),
        ("first_idr_ipcm", ctypes.c_uint32), ("part_mix", ctypes.c_uint32),
        ("crop", ctypes.c_uint32), ("multi_slice_params", ctypes.c_uint32), ("dpb_stress", ctypes.c_uint32), ("fmo_type", ctypes.c_uint32), ("fmo_groups", ctypes.c_uint32),
    ]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libh264writer.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = ctypes.CDLL(path)
        _lib.h264w_default_params.argtypes = [ctypes.POINTER(WriterParams), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        _lib.h264w_bound.argtypes = [ctypes.POINTER(WriterParams)]
        _lib.h264w_bound.restype = ctypes.c_size_t
        _lib.h264w_generate.argtypes = [ctypes.POINTER(WriterParams), ctypes.c_void_p, ctypes.c_size_t]
        _lib.h264w_generate.restype = ctypes.c_size_t
    return _lib


def default_params(width_mbs, height_mbs, n_frames, **overrides):
    lib = _load()
    p = WriterParams()
    lib.h264w_default_params(ctypes.byref(p), width_mbs, height_mbs, n_frames)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def generate(params) -> bytes:
    """Serialise the stream described by `params` to Annex-B bytes."""
    lib = _load()
    cap = lib.h264w_bound(ctypes.byref(params))
    buf = ctypes.create_string_buffer(cap)
    n = lib.h264w_generate(ctypes.byref(params), buf, cap)
    if n == 0:
        raise RuntimeError("h264w_generate failed (bad parameters or buffer too small)")
    return buf.raw[:n]


def synth(width_mbs, height_mbs, n_frames, **overrides) -> bytes:
    return generate(default_params(width_mbs, height_mbs, n_frames, **overrides))
