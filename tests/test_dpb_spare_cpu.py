"""The spare frame slot (h264_dpb_rotate_spare): tests/native/dpb_spare_check.c drives the product's DPB through IPPP
sequences and checks that a picture is never decoded into a slot the DPB still holds, nor into the slot of the picture
before it — the property that lets the engine launch a picture while the copy-out of the one before last still runs."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spare_slot_rotation(tmp_path):
    exe = str(tmp_path / "dpb_spare_check")
    c = os.path.join(ROOT, "broadway_b200", "csrc")
    subprocess.run(["gcc", "-O1", "-g", "-I" + os.path.join(ROOT, "include"), "-I" + c,
                    os.path.join(ROOT, "tests", "native", "dpb_spare_check.c"), os.path.join(c, "h264_dpb.c"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
