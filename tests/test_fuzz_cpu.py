"""CPU robustness tests of the host parser (the same C sources the product library links): corrupted
and truncated streams must be rejected or decoded partially — never crash, hang or overrun.  Runs the
parser under the oracle backend (oracle/cpuchkdec), which also exercises the reconstruction
restatement on whatever records survive."""
import os
import random
import subprocess
import tempfile

import cases
import util


def _run(data):
    with tempfile.NamedTemporaryFile(suffix=".264", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        r = subprocess.run([util.CPUCHK, "-m", path], capture_output=True, text=True, timeout=60)
    finally:
        os.remove(path)
    return r


def test_bit_flips_never_crash():
    rng = random.Random(2026)
    base = [cases.make_stream(c) for c in cases.SMALL[:6]]
    for i in range(60):
        data = bytearray(rng.choice(base))
        for _ in range(rng.randrange(1, 8)):
            pos = rng.randrange(len(data))
            data[pos] ^= 1 << rng.randrange(8)
        r = _run(bytes(data))
        assert r.returncode >= 0, "decoder died with signal %d on mutation %d" % (-r.returncode, i)


def test_truncation_and_garbage_never_crash():
    rng = random.Random(7)
    data = cases.make_stream(cases.SMALL[0])
    for cut in [0, 1, 3, 4, 5, 17, 100, len(data) // 3, len(data) - 1]:
        assert _run(data[:cut]).returncode >= 0
    for i in range(10):
        junk = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 4000)))
        assert _run(junk).returncode >= 0
        assert _run(b"\x00\x00\x00\x01" + junk).returncode >= 0
        assert _run(data[:200] + junk + data[200:]).returncode >= 0


def test_zero_length_nals_and_repeated_headers():
    data = cases.make_stream(cases.SMALL[0])
    nals = [n for n in data.split(b"\x00\x00\x00\x01") if n]
    sc = b"\x00\x00\x00\x01"
    doubled = sc + nals[0] + sc + nals[1] + sc + nals[0] + sc + nals[1] + b"".join(sc + n for n in nals[2:])
    r = _run(doubled)
    assert r.returncode == 0
    want, _ = util.oracle_md5(data)
    got = [l.split()[2] for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert got == want                                    # re-sent identical parameter sets change nothing
    assert _run(sc + sc + sc + data).returncode >= 0
