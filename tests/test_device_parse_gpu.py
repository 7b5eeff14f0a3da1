"""GPU parity tests of the DEVICE-PARSE path (run on the B200 box with -m gpu): slice data parsed by CUDA kernel Kp
(broadway_b200/csrc/kp_core.h, kp_parse.cuh) instead of the host parser, everything through the C ABI of
libh264b200.so.  Checkers: the committed golden MD5s of the unmodified reference, the oracle (CPU restatement run
live), and — for the records Kp writes into HBM — the HOST parser's records for the same stream, byte for byte
("device records == host records" on every case of tests/cases.py)."""
import json
import os
import random

import pytest

import cases
import util
from broadway_b200 import bitstream, capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "loss.json")))
DEV = capi.ENGINE_DEVICE_PARSE


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    capi.require_gpu()


@pytest.fixture(scope="module")
def sync_engine():
    """not batched: every picture is parsed (Kp) and reconstructed as soon as its access unit ends"""
    with capi.Engine(flags=DEV) as eng:
        yield eng


@pytest.fixture(scope="module")
def parse_only_engine():
    """Kp only (no K1..K4): the coefficient slots still hold the levels Kp wrote (K1 transforms them in place)"""
    with capi.Engine(flags=DEV | capi.ENGINE_NO_RECON) as eng:
        yield eng


@pytest.mark.parametrize("case", cases.SMALL, ids=[c[0] for c in cases.SMALL])
def test_device_parse_matches_golden_and_host_records(case, golden, sync_engine, parse_only_engine):
    data = cases.make_stream(case)
    got, info = capi.decode_on_engine(sync_engine, data)
    assert info["device_parse"] == 1
    assert info["err_mbs"] == 0
    assert got == golden[case[0]]["frame_md5"]
    # what Kp left in HBM against the host parser's records and coefficient slots
    _, _, parses = capi.decode_on_engine(parse_only_engine, data, fetch_parse=True)
    host = util.capture_records(util.cpuchk_lib(), data, False)
    assert len(parses) == len(host) == case[3]
    n_mbs = case[1] * case[2]
    for k, ((mbs, coef, res), (hmbs, hcoef)) in enumerate(zip(parses, host)):
        assert res[0] * 32 == len(hcoef), "picture %d: %d slots on the device, %d on the host" % (k, res[0], len(hcoef) // 32)
        assert coef == hcoef, "coefficient slots of picture %d differ" % k
        assert util.canonical_records(mbs, n_mbs) == hmbs, "records of picture %d differ" % k
        assert res[8] == 0 and res[10] == n_mbs          # err_mbs, decoded_mbs


@pytest.mark.parametrize("case", cases.FULL, ids=[c[0] for c in cases.FULL])
def test_device_parse_full_size_matches_reference_golden(case, golden, sync_engine):
    got, info = capi.decode_on_engine(sync_engine, cases.make_stream(case))
    assert (info["width"], info["height"]) == (16 * case[1], 16 * case[2])
    assert got == golden[case[0]]["frame_md5"]


@pytest.mark.parametrize("lc", cases.LOSS, ids=[c[0] for c in cases.LOSS])
def test_device_parse_concealment_matches_reference_golden(lc, sync_engine):
    g = LOSS_GOLDEN[lc[0]]
    got, info = capi.decode_on_engine(sync_engine, cases.make_loss_stream(lc))
    assert got == g["frame_md5"]
    assert info["err_mbs"] == g["err_mbs"]


def test_device_parse_batched_streams_match_golden(golden):
    """Every small case at once through the look-ahead pipeline: pictures of different sizes, slice structures and DPB
    behaviour share Kp launches and reconstruction rounds."""
    sel = cases.SMALL
    streams = [cases.make_stream(c) for c in sel]
    for threads in (1, 4):
        with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
            md5s, rs = eng.decode_streams_md5(streams, threads=threads)
            assert rs.failed_streams == 0 and rs.err_mbs == 0
            for c, m in zip(sel, md5s):
                assert m == golden[c[0]]["frame_md5"], (c[0], threads)
            st = eng.stats()
            assert st["pictures"] == sum(c[3] for c in sel)
            assert st["kp_pictures"] == st["pictures"] and 0 < st["kp_launches"] < st["pictures"] // 4
            assert eng.error_flags() == 0


def test_device_parse_batched_reports_concealed_macroblocks():
    sel = cases.LOSS
    streams = [cases.make_loss_stream(c) for c in sel]
    errs = [0] * len(sel)
    md5 = [dict() for _ in sel]

    def on_picture(stream, index, ptr, w, h, pic_id, err):
        errs[stream] += err
        md5[stream][index] = capi.frame_md5(ptr, w * h * 3 // 2)
    with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
        eng.decode_streams(streams, 3, on_picture)
    for c, m, e in zip(sel, md5, errs):
        g = LOSS_GOLDEN[c[0]]
        assert [m[i] for i in sorted(m)] == g["frame_md5"], c[0]
        assert e == g["err_mbs"], c[0]


def test_device_parse_bit_errors_equal_the_oracle(sync_engine):
    """Corrupted slices: Kp must notice the damage at the same macroblock as the host parser (it is the same algorithm),
    give the same macroblocks back and conceal the same way — whatever comes out equals the CPU oracle."""
    rng = random.Random(303)
    bases = [cases.make_stream(c) for c in cases.SMALL[:14]]
    for _ in range(25):
        data = bytearray(rng.choice(bases))
        for _ in range(rng.randrange(1, 4)):
            data[rng.randrange(60, len(data))] ^= 1 << rng.randrange(8)
        want, sw = util.oracle_md5(bytes(data))
        got, info = capi.decode_on_engine(sync_engine, bytes(data))
        assert got == want and info["err_mbs"] == sw["err_mbs"]


def test_device_parse_1080p_window_equals_host_parse():
    """Size-independent property at BASELINE.json's size: 1080p streams through the look-ahead pipeline (device parse)
    give exactly the frames of the host-parse engine, and the first stream those of the oracle."""
    streams = [bitstream.synth(120, 68, 6, seed=777 + i) for i in range(6)]
    with capi.Engine(flags=capi.ENGINE_BATCHED) as eng:
        host, _ = eng.decode_streams_md5(streams, threads=4)
    with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
        dev, rs = eng.decode_streams_md5(streams, threads=4)
        assert rs.err_mbs == 0
    assert dev == host
    want, _ = util.oracle_md5(streams[0])
    assert dev[0] == want


def test_device_parse_resident_replay_reproduces_the_pictures():
    """bench.py's resident leg in device-parse mode replays the tape of Kp launches and reconstruction rounds from
    HBM-resident slice blocks: it must rebuild the same frames."""
    streams = [bitstream.synth(20, 12, 7, seed=300 + i, p_intra_permille=100) for i in range(6)]
    with capi.Engine(flags=capi.ENGINE_BATCHED | capi.ENGINE_RETAIN | DEV) as eng:
        md5s, _ = eng.decode_streams_md5(streams, threads=2)
        for s, m in zip(streams, md5s):
            assert m == util.oracle_md5(s)[0]
        assert eng.check_resident() == 0
        n = eng.replay(reps=3, time_kernels=True)
        assert n == 3 * 42
        assert eng.check_resident() == 0
        kt = eng.kernel_times()
        assert kt["kp_parse"]["launches"] >= 3 and kt["kp_parse"]["ms"] > 0 and kt["kp_parse"]["bytes"] > 0
        assert kt["k4_deblock"]["ms"] > 0 and kt["k2_inter"]["bytes"] > 0


def test_device_parse_thread_count_and_window_do_not_change_results(monkeypatch):
    streams = [cases.make_stream(c) for c in cases.SMALL[:10]]
    ref = None
    for threads, window in ((1, "1"), (3, "4"), (10, "16")):
        monkeypatch.setenv("H264B200_WINDOW", window)
        with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
            md5s, _ = eng.decode_streams_md5(streams, threads=threads)
        if ref is None:
            ref = md5s
        assert md5s == ref


@pytest.mark.parametrize("threads,host", [(1, 3), (4, 9), (8, 99)])
def test_host_share_of_a_device_parse_run_matches_golden(golden, monkeypatch, threads, host):
    """h264b200DecodeStreams gives `host` streams of a device-parse run to the worker threads' parser (h264b200SetHostParse):
    host- and device-parsed pictures share Kp-less and Kp-fed rounds of one engine, and every picture equals the reference
    golden whoever parsed it."""
    monkeypatch.setenv("H264B200_HOST_STREAMS", str(host))
    sel = cases.SMALL
    streams = [cases.make_stream(c) for c in sel]
    with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
        md5s, rs = eng.decode_streams_md5(streams, threads=threads)
        assert rs.failed_streams == 0 and rs.err_mbs == 0
        assert rs.host_streams == min(host, len(sel))
        for c, m in zip(sel, md5s):
            assert m == golden[c[0]]["frame_md5"], (c[0], threads, host)
        st = eng.stats()
        assert st["pictures"] == sum(c[3] for c in sel)
        assert st["kp_pictures"] < st["pictures"]
        assert eng.error_flags() == 0


def test_host_share_1080p_equals_all_device(monkeypatch):
    from broadway_b200 import bitstream
    streams = [bitstream.synth(120, 68, 6, seed=40 + i) for i in range(6)]
    out = []
    for host in ("0", "3"):
        monkeypatch.setenv("H264B200_HOST_STREAMS", host)
        with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
            md5s, rs = eng.decode_streams_md5(streams, threads=3)
            assert rs.failed_streams == 0 and rs.host_streams == int(host)
        out.append(md5s)
    assert out[0] == out[1]


def test_free_running_pipeline_is_repeatable(golden):
    """The free-running pipeline has three kinds of threads — workers that poll picture states without the engine mutex, the
    scheduling thread, the GPU — and the first version handed a worker a frame whose copy-out had not been issued yet
    (the launch was published before its event was recorded).  Twenty runs at several thread counts, every picture
    against the reference golden each time."""
    sel = cases.SMALL
    streams = [cases.make_stream(c) for c in sel]
    with capi.Engine(flags=capi.ENGINE_BATCHED | DEV) as eng:
        for rep in range(20):
            threads = (2, 3, 4, 8, 16)[rep % 5]
            md5s, rs = eng.decode_streams_md5(streams, threads=threads)
            assert rs.failed_streams == 0 and rs.err_mbs == 0
            for c, m in zip(sel, md5s):
                assert m == golden[c[0]]["frame_md5"], (c[0], threads, rep)
