"""Build recipe of the native libraries (called by __graft_entry__.build()).

  broadway_b200/libh264b200.so   the product: host decoder (C) + CUDA engine, sm_100a only
  broadway_b200/libh264writer.so synthetic bitstream writer (test/bench input generator)
  broadway_b200/bin/b200dec      tools/swdec_cli.c linked against libh264b200.so
  oracle/...                     test infrastructure (oracle/Makefile): the CPU
                                 restatement and, when /root/reference is mounted,
                                 the unmodified reference decoder (oracle/_ref/)

Everything is built IN-TREE so the shared objects travel to the GPU box.
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "broadway_b200")
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(ROOT, "build")

HOST_C = ["h264_decoder.c", "h264_params.c", "h264_dpb.c", "h264_slice.c", "h264_cavlc.c",
          "h264_swdec.c", "h264_runner.c", "h264_mp4.c", "h264_shim.c"]
CUDA = ["h264_engine.cu"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built")


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs += [os.path.join(INC, f) for f in os.listdir(INC)]
    return hs


def build_product(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    log = []
    for f in HOST_C:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f + ".o")
        if force or _newer(obj, [src] + hdrs):
            log.append(_run(["gcc", "-O3", "-g", "-fPIC", "-Wall", "-Wextra", "-pthread", "-I" + INC, "-I" + CSRC, "-c", src, "-o", obj]))
        objs.append(obj)
    for f in CUDA:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f + ".o")
        if force or _newer(obj, [src] + hdrs):
            log.append(_run([_nvcc()] + NVCC_ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
                                                     "-I" + INC, "-I" + CSRC, "-c", src, "-o", obj]))
        objs.append(obj)
    lib = os.path.join(PKG, "libh264b200.so")
    if force or _newer(lib, objs):
        log.append(_run([_nvcc()] + NVCC_ARCH + ["-shared", "-o", lib] + objs + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]))
    bindir = os.path.join(PKG, "bin")
    os.makedirs(bindir, exist_ok=True)
    cli = os.path.join(bindir, "b200dec")
    cli_src = os.path.join(ROOT, "tools", "swdec_cli.c")
    if force or _newer(cli, [cli_src, lib]):
        log.append(_run(["gcc", "-O2", "-DUSE_B200", "-I" + INC, cli_src, "-o", cli, "-L" + PKG, "-lh264b200", "-Wl,-rpath,$ORIGIN/.."]))
    if verbose:
        print("\n".join(log))
    return lib


def build_writer(force=False):
    lib = os.path.join(PKG, "libh264writer.so")
    src = os.path.join(CSRC, "h264_writer.c")
    if force or _newer(lib, [src] + _headers()):
        _run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-Wextra", "-I" + INC, "-I" + CSRC, src, "-o", lib])
    return lib


def build_oracle():
    """Test infrastructure: CPU restatement always; the reference itself only where its sources are mounted."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False, verbose=False):
    build_writer(force)
    build_product(force, verbose)
    build_oracle()
