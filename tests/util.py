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
