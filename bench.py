#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: 1080p frames/s of bit-exact H.264 Baseline decode.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2], "synthetic 1080p Baseline IPPP stream
(random MVs incl. quarter-pel, deblocking on)" — configs[1] (Player/tree.mp4) is absent from the
reference mount.  Per GPU: S independent 1920x1088 streams of F pictures (1 IDR + F-1 P; all
partition shapes, quarter-pel vectors, ~30 % coded 4x4 blocks, in-loop deblocking on) from the
in-repo writer.  One STEP decodes all S*F pictures: F reconstruction rounds of S pictures per kernel
family, fed by Kp launches over the look-ahead window of every stream.  Weak scaling: every rank gets its own S streams.

Own arm prints
  value       frames/s with the slices already resident in HBM: the retained tape of one decode — every launch of
              kernel Kp (CAVLC / macroblock-layer parse on the device) and every reconstruction round (K1 transform,
              K2 inter, K3 intra, K4 deblock) with the dependencies of the live run — replayed without host<->device
              copies; CUDA events on the engine's compute stream; the replay is checked to rebuild the very same frames
  e2e         frames/s through the C ABI (h264b200DecodeStreams -> h264bsdDecode per NAL) from HOST Annex-B bytes to
              HOST I420 frames: NAL scan + slice headers + DPB on the host threads, H2D of the slices, Kp, K1..K4, D2H
              of every frame into pinned memory — all inside the timed region (--parse host: round 1's path, slice
              data parsed on the host cores and records uploaded)
  roofline    the kernel family with the largest share of device time, algorithmic bytes
  cpu_baseline the UNMODIFIED reference decoder (oracle/_ref/refdec, gcc -O3), one process per host core
Reference arm (--impl reference) times that same reference build on the same streams.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH_MBS, HEIGHT_MBS = 120, 68          # 1920x1088 coded size of 1080p (default workload)
WORKLOADS = {
    # name: (width_mbs, height_mbs, writer overrides, description)
    "1080p_ippp": (120, 68, {}, "synthetic 1080p Baseline IPPP, random MVs incl. quarter-pel, ~30% coded blocks, deblocking on (BASELINE.json configs[2]; configs[1] tree.mp4 absent)"),
    "1080p_intra": (120, 68, {"intra_only": 1}, "synthetic 1080p intra-only (I16x16/I4x4 mix) stressing the intra and deblock wavefronts (BASELINE.json configs[3])"),
    "4k_ippp": (240, 135, {"level_idc": 51}, "synthetic 4K (3840x2160) Baseline IPPP multi-stream (BASELINE.json configs[4])"),
    "4k_gop": (240, 135, {"level_idc": 51, "idr_period": 4}, "synthetic 4K (3840x2160) Baseline, IDR every 4 pictures, IDR-bounded GOP segments sharded over the GPUs (BASELINE.json configs[4])"),
    # not BASELINE.json configurations: the same 1080p IPPP structure at streaming bitrates, to show how the host-parse
    # bound of e2e moves with the bitrate; ~54 Mbit/s at 30 frames/s for the default workload
    "1080p_ippp_17mbps": (120, 68, {"coded_blk_permille": 40, "p_skip_permille": 400}, "synthetic 1080p Baseline IPPP at ~17 Mbit/s (4% coded blocks, 40% P_Skip); illustration, not a BASELINE.json config"),
    "1080p_ippp_6mbps": (120, 68, {"coded_blk_permille": 20, "p_skip_permille": 600, "part_mix": 0}, "synthetic 1080p Baseline IPPP at ~6 Mbit/s (2% coded blocks, 60% P_Skip, 16x16 partitions); illustration, not a BASELINE.json config"),
}
_WL = {"name": "1080p_ippp"}
DEFAULT_STREAMS, DEFAULT_FRAMES, DISTINCT = 256, 64, 16
REFDEC = os.path.join(ROOT, "oracle", "_ref", "refdec")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_streams(n_streams, n_frames, rank):
    """S streams per rank from DISTINCT distinct seeds (the writer produces ~16 1080p pictures per second)."""
    from concurrent.futures import ThreadPoolExecutor
    from broadway_b200 import bitstream
    n_distinct = min(DISTINCT, n_streams)

    w, h, kw, _ = WORKLOADS[_WL["name"]]

    def gen(i):
        return bitstream.synth(w, h, n_frames, seed=1234 + 1000 * rank + i, **kw)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        base = list(ex.map(gen, range(n_distinct)))
    return [base[i % n_distinct] for i in range(n_streams)]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full summary, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------- reference (CPU) timing
def run_reference(streams, n_procs, reps):
    """One refdec process per core (taskset-pinned), process i decodes streams[i % len] `reps` times.
    Returns (frames, wall seconds)."""
    tmp = tempfile.mkdtemp(prefix="h264ref_")
    paths = []
    for i, s in enumerate(streams[:n_procs]):
        p = os.path.join(tmp, "s%d.264" % i)
        open(p, "wb").write(s)
        paths.append(p)
    have_taskset = subprocess.run(["which", "taskset"], capture_output=True).returncode == 0
    cpus = sorted(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    procs = []
    for i in range(n_procs):
        cmd = [REFDEC, "-r", str(reps), paths[i % len(paths)]]
        if have_taskset:
            cmd = ["taskset", "-c", str(cpus[i % len(cpus)])] + cmd
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    frames = 0
    for p in procs:
        out = p.communicate()[0]
        frames += json.loads(out.splitlines()[-1])["frames"]
    dt = time.perf_counter() - t0
    for p in paths:
        os.remove(p)
    os.rmdir(tmp)
    return frames, dt


def reference_arm(args, rank, world):
    if rank != 0:
        return
    if not os.path.exists(REFDEC):
        emit(({"impl": "reference", "unavailable": "oracle/_ref/refdec not built (reference sources absent at build time)"}))
        return
    cores = len(os.sched_getaffinity(0))
    streams = make_streams(min(cores, DISTINCT), args.frames, 0)
    for _ in range(args.warmup):
        run_reference(streams, cores, 1)
    frames = 0
    t = 0.0
    for _ in range(args.steps):
        f, dt = run_reference(streams, cores, 1)
        frames += f
        t += dt
    fps = frames / t
    sample = "%d refdec processes (one per host core), each decoding one %d-picture 1080p stream per step" % (cores, args.frames)
    emit(({
        "impl": "reference", "metric": "1080p frames/sec (bit-exact H.264 Baseline decode)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "width": 16 * WORKLOADS[args.workload][0], "height": 16 * WORKLOADS[args.workload][1],
                   "frames_per_stream": args.frames, "streams_per_step": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- own arm
def _cpulist(text):
    out = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def bind_to_gpu_node(local_rank, world):
    """Multi-GPU boxes have more than one NUMA node: keep this rank's threads — and with them the pinned frame mirrors they
    first touch — on the host cores next to its GPU (sysfs local_cpulist of the PCI device), sharing a node's cores evenly
    between the ranks whose GPUs hang off it.  A no-op when sysfs says nothing (one node, virtual topology).  Returns a
    note for the JSON line."""
    try:
        import torch
        allowed = os.sched_getaffinity(0)
        local = {}
        for lr in range(world):
            bus = torch.cuda.get_device_properties(lr).pci_bus_id if hasattr(torch.cuda.get_device_properties(lr), "pci_bus_id") else None
            dom = getattr(torch.cuda.get_device_properties(lr), "pci_domain_id", 0)
            dev = getattr(torch.cuda.get_device_properties(lr), "pci_device_id", 0)
            if bus is None:
                return "no pci id"
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
            local[lr] = frozenset(_cpulist(open(path).read()) & allowed)
        mine = local[local_rank]
        if not mine or mine == frozenset(allowed) and world == 1:
            return "single node"
        peers = sorted(lr for lr in range(world) if local[lr] == mine)
        cpus = sorted(mine)
        share = cpus[peers.index(local_rank)::len(peers)]
        if not share:
            return "no share"
        os.sched_setaffinity(0, share)
        return "rank bound to %d of the %d cores local to its GPU" % (len(share), len(cpus))
    except Exception as e:                                   # topology not exposed: leave the affinity alone
        return "unbound (%s)" % type(e).__name__


def own_arm(args, rank, local_rank, world):
    import torch
    from broadway_b200 import capi
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    capi.require_gpu()
    cores = len(os.sched_getaffinity(0))
    threads = args.threads or max(1, cores // world)
    streams = make_streams(args.streams, args.frames, rank)
    numa_note = bind_to_gpu_node(local_rank, world) if world > 1 else "single GPU"
    frames_per_step = args.streams * args.frames
    pflag = capi.ENGINE_DEVICE_PARSE if args.parse == "device" else 0
    log("[rank %d] %d streams x %d pictures, %.1f MB of Annex-B, %d parser threads" % (rank, args.streams, args.frames, sum(map(len, streams)) / 1e6, threads))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity gate (untimed): every picture of the first streams against the committed per-frame MD5s
    # of the UNMODIFIED reference (tests/golden/bench_streams.json, made by tools/make_golden.py)
    check = {}
    gpath = os.path.join(ROOT, "tests", "golden", "bench_streams.json")
    if rank == 0 and not args.no_check and os.path.exists(gpath) and args.workload == "1080p_ippp":
        import hashlib
        g = json.load(open(gpath))
        if g["frames"] == args.frames:
            n_chk = min(len(g["streams"]), len(streams))
            with capi.Engine(local_rank, capi.ENGINE_BATCHED | pflag) as eng:
                got, _ = eng.decode_streams_md5(streams[:n_chk], threads=n_chk)
            for i in range(n_chk):
                if hashlib.md5(streams[i]).hexdigest() != g["streams"][i]["stream_md5"]:
                    raise SystemExit("bench stream %d is not the stream the golden was made from" % i)
                if got[i] != g["streams"][i]["frame_md5"]:
                    raise SystemExit("PARITY FAILURE on bench stream %d against the reference golden" % i)
            check = {"streams": n_chk, "pictures": n_chk * args.frames, "against": "reference golden MD5 (tests/golden/bench_streams.json)", "md5_exact": True}
            log("[rank 0] parity gate: %d pictures MD5-exact vs the reference" % check["pictures"])

    # ---- end-to-end leg: host Annex-B -> host I420 frames through the C ABI.  The K timed steps are ONE streaming call:
    # every stream is its 64-picture step repeated K times back to back (each repetition starts with its parameter sets
    # and an IDR picture, so the concatenation is a valid stream that decodes to K times the same pictures), i.e. the
    # decoder sees 256 streams of K x 64 pictures and its pipeline (host scan -> Kp -> K1..K4 -> copy-out) is filled and
    # drained once per timed region, not once per step — a per-step call would measure the 0.3 s pipeline latency
    # (one picture is ~0.29 s of serial parsing on one warp) K times over.
    def repeated(k):
        cache = {}
        return [cache.setdefault(id(b), b * k) for b in streams]
    eng = capi.Engine(local_rank, capi.ENGINE_BATCHED | pflag)
    if not args.skip_e2e and args.warmup:
        eng.decode_streams(repeated(args.warmup), threads)
    timed = repeated(args.steps)
    barrier()
    s0 = eng.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    parse_s = wait_s = 0.0
    host_streams = 0
    if not args.skip_e2e:
        rs = eng.decode_streams(timed, threads)
        assert rs.pictures == frames_per_step * args.steps and rs.err_mbs == 0, (rs.pictures, rs.err_mbs)
        host_streams = rs.host_streams
        parse_s += rs.parse_seconds
        wait_s += rs.wait_seconds
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    s1 = eng.stats()
    eng.close()
    del timed
    e2e_fps = world * frames_per_step * args.steps / e2e_s
    h2d = (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps
    d2h = (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps
    launches_e2e = s1["kernel_launches"] - s0["kernel_launches"]

    if args.e2e_only:
        if rank == 0:
            sampler.stop()
            emit({"e2e_only": True, "e2e_fps": e2e_fps, "threads": threads, "streams": args.streams, "host_streams": host_streams, "ms_per_step": 1000.0 * e2e_s / args.steps,
                  "parse_core_s": parse_s / args.steps, "wait_s": wait_s / args.steps})
        return

    # ---- resident leg: retain every launch of one decode in HBM (Kp launches and reconstruction rounds), then replay the
    # device work alone.  Every stream is parsed by kernel Kp here (no host share): `value` is the GPU doing the whole job.
    eng = capi.Engine(local_rank, capi.ENGINE_BATCHED | capi.ENGINE_RETAIN | pflag)
    keep = os.environ.get("H264B200_HOST_STREAMS")
    os.environ["H264B200_HOST_STREAMS"] = "0"
    eng.decode_streams(streams, threads)
    if keep is None:
        del os.environ["H264B200_HOST_STREAMS"]
    else:
        os.environ["H264B200_HOST_STREAMS"] = keep
    eng.sync()
    assert eng.check_resident() == 0
    for _ in range(args.warmup):
        eng.replay(1, False)
    eng.sync()
    eng.kernel_times(reset=True)
    barrier()
    s0 = eng.stats()
    n = eng.replay(args.steps, True)
    dev_ms = eng.replay_ms()
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    s1 = eng.stats()
    assert n == frames_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    kt = eng.kernel_times()
    bad = eng.check_resident()
    assert bad == 0, "replay did not reproduce the decoded frames (%d slots differ)" % bad
    err = eng.error_flags()
    eng.close()
    value = world * frames_per_step * args.steps / (dev_ms / 1000.0)
    launches = s1["kernel_launches"] - s0["kernel_launches"]

    # ---- roofline of the dominant kernel family
    peak, peak_src = measured_peak()
    dom = max(kt, key=lambda k: kt[k]["ms"])
    kinfo = {}
    total_ms = sum(v["ms"] for v in kt.values()) or 1.0
    for k, v in kt.items():
        if v["launches"]:
            kinfo[k] = {"ms_per_launch": v["ms"] / v["launches"], "GBps": v["bytes"] / v["ms"] / 1e6 if v["ms"] else None,
                        "frac_of_peak": (v["bytes"] / v["ms"] / 1e6 / peak) if v["ms"] else None, "share_of_step": v["ms"] / total_ms,
                        "launches": v["launches"], "algorithmic_bytes_per_launch": v["bytes"] // v["launches"]}
    ach = kinfo[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(dom), "peak_source": peak_src, "kernels": kinfo}

    # ---- CPU baseline: the unmodified reference on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if os.path.exists(REFDEC):
            reps = max(1, int(round(200.0 / args.frames)))           # ~200 pictures per core: 10-15 s
            f, dt = run_reference(streams[:min(cores, DISTINCT)], cores, reps)
            cpu = {"value": f / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
                   "sample": "%d refdec processes (unmodified reference, gcc -O3, one per host core) x %d pictures of the same 1080p streams" % (cores, args.frames * reps)}
        else:
            cpu = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/refdec missing"}

    if rank == 0:
        emit(({
            "metric": "%s frames/sec (bit-exact H.264 Baseline decode)" % ("4K" if args.workload.startswith("4k") else "1080p"), "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][3],
                       "width": 16 * WORKLOADS[args.workload][0], "height": 16 * WORKLOADS[args.workload][1], "streams_per_gpu": args.streams, "frames_per_stream": args.frames,
                       "frames_per_step_per_gpu": frames_per_step, "parser_threads_per_gpu": threads, "host_cores": cores, "slice_data_parse": args.parse, "numa": numa_note,
                       "l2": "inputs larger than L2 (per step: %.0f MB of frame pools + records per GPU)" % (args.streams * 2 * WORKLOADS[args.workload][0] * WORKLOADS[args.workload][1] * 384 / 1e6 + h2d / 1e6)},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000.0 * e2e_s / args.steps, "timing": "wall clock between barrier+synchronize pairs around ONE streaming call over the K steps (each stream = its step repeated K times), max over ranks",
                    "scaling_note": "weak scaling: every rank decodes its own %d streams on its own GPU with %d host threads, ranks never interact; compare this e2e value across N, not `value`" % (args.streams, threads),
                    "host_parsed_streams": host_streams, "slice_data_parse": ("kernel Kp for %d streams, the %d worker threads' own parser for %d (h264b200DecodeStreams' host share)" % (args.streams - host_streams, threads, host_streams)) if args.parse == "device" else "host",
                    "host_parse_core_seconds_per_step": parse_s / args.steps, "host_wait_seconds_per_step": wait_s / args.steps,
                    "kernel_launches": launches_e2e},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "parity": check, "device_error_flags": err,
        }))
    if dist:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------- GOP-sharded arm (BASELINE configs[4])
GOP_DISTINCT, GOP_FRAMES, GOP_REPLICAS = 2, 128, 4          # 2 distinct 4K streams x 128 pictures (32 GOPs of 4), each 4 times: 256 units


def gop_streams():
    from concurrent.futures import ThreadPoolExecutor
    from broadway_b200 import bitstream
    w, h, kw, _ = WORKLOADS["4k_gop"]
    with ThreadPoolExecutor(max_workers=GOP_DISTINCT) as ex:
        return list(ex.map(lambda i: bitstream.synth(w, h, GOP_FRAMES, seed=4321 + i, **kw), range(GOP_DISTINCT)))


def gop_arm(args, rank, local_rank, world):
    """STRONG scaling: a fixed set of 4K streams is cut at its IDR access units (h264b200SplitGops), the segments are
    assigned to the ranks by size (broadway_b200/shard.py), every rank decodes its segments as independent streams on
    its own GPU, the per-picture digests are gathered (all_gather_object: the only cross-rank traffic) and compared
    with the committed digests of the SEQUENTIAL decode of the whole streams by the unmodified reference
    (tests/golden/gop_4k.json).  The timed region is the decode alone (host Annex-B segments -> host I420 frames)."""
    import hashlib
    import torch
    from broadway_b200 import capi, shard
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    capi.require_gpu()
    cores = len(os.sched_getaffinity(0))
    threads = args.threads or max(1, cores // world)
    base = gop_streams()
    segs_of = [capi.split_gops(b) for b in base]
    units, owner = [], []
    for rep in range(GOP_REPLICAS):
        for i, segs in enumerate(segs_of):
            units += segs
            owner += [(i, k) for k in range(len(segs))]
    plan = shard.assign([len(u) for u in units], world)
    mine = [units[i] for i in plan[rank]]
    my_pics = sum(GOP_FRAMES // len(segs_of[owner[i][0]]) for i in plan[rank])
    total_pics = GOP_REPLICAS * GOP_DISTINCT * GOP_FRAMES
    pflag = capi.ENGINE_DEVICE_PARSE if args.parse == "device" else 0
    log("[rank %d] %d of %d GOP segments, %d of %d pictures, %d threads" % (rank, len(mine), len(units), my_pics, total_pics, threads))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity (untimed): digests of every segment, gathered, against the sequential golden
    check = {}
    gpath = os.path.join(ROOT, "tests", "golden", "gop_4k.json")
    if not args.no_check and os.path.exists(gpath):
        g = json.load(open(gpath))

        def decode_fn(us):
            with capi.Engine(local_rank, capi.ENGINE_BATCHED | pflag) as eng:
                md5s, _ = eng.decode_streams_md5(us, threads=threads)
            return md5s
        res = shard.decode_sharded(units, decode_fn, rank, world, dist)
        if rank == 0:
            for i, b in enumerate(base):
                if hashlib.md5(b).hexdigest() != g["streams"][i]["stream_md5"]:
                    raise SystemExit("4K stream %d is not the stream the golden was made from" % i)
            for rep in range(GOP_REPLICAS):
                for i in range(GOP_DISTINCT):
                    got = [m for (u, (si, k)) in zip(res, owner) if si == i for m in u][rep * GOP_FRAMES:(rep + 1) * GOP_FRAMES]
                    if got != g["streams"][i]["frame_md5"]:
                        raise SystemExit("PARITY FAILURE: GOP-sharded decode of 4K stream %d differs from the sequential reference decode" % i)
            check = {"pictures": total_pics, "units": len(units), "against": "sequential decode by the unmodified reference (tests/golden/gop_4k.json)", "md5_exact": True}
            log("[rank 0] parity: %d pictures over %d ranks MD5-exact vs the sequential reference decode" % (total_pics, world))

    eng = capi.Engine(local_rank, capi.ENGINE_BATCHED | pflag)
    for _ in range(args.warmup):
        if mine:
            eng.decode_streams(mine, threads)
    barrier()
    s0 = eng.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if mine:
            rs = eng.decode_streams(mine, threads)
            assert rs.pictures == my_pics and rs.err_mbs == 0, (rs.pictures, my_pics, rs.err_mbs)
    barrier()
    dt = time.perf_counter() - t0
    if dist:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    s1 = eng.stats()
    eng.close()
    clocks = sampler.stop() if rank == 0 else None
    fps = total_pics * args.steps / dt
    if rank == 0:
        emit({"metric": "4K frames/sec (bit-exact H.264 Baseline decode, IDR-bounded GOP segments sharded over the GPUs)", "value": fps, "unit": "frames/s",
              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
              "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
              "config": {"workload": WORKLOADS["4k_gop"][3], "width": 3840, "height": 2160, "streams": GOP_DISTINCT * GOP_REPLICAS, "frames_per_stream": GOP_FRAMES,
                         "gop_segments": len(units), "segments_on_rank0": len(mine), "parser_threads_per_gpu": threads, "host_cores": cores, "slice_data_parse": args.parse,
                         "l2": "inputs larger than L2"},
              "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps,
                      "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps, "timing": "wall clock between barrier+synchronize pairs, max over ranks; value IS the end-to-end number in this mode (rank 0's byte counts)"},
              "gpu_launches": s1["kernel_launches"] - s0["kernel_launches"], "clocks": clocks, "parity": check})
    if dist:
        dist.destroy_process_group()


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else printed by libraries (NCCL banner ...)
    was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=DEFAULT_STREAMS, help="independent 1080p streams per GPU")
    ap.add_argument("--frames", type=int, default=DEFAULT_FRAMES, help="pictures per stream (1 IDR + P)")
    ap.add_argument("--threads", type=int, default=0, help="parser threads per GPU (0: host cores / ranks)")
    ap.add_argument("--workload", default="1080p_ippp", choices=sorted(WORKLOADS), help="default: the configuration BASELINE.json's metric is quoted on")
    ap.add_argument("--parse", default="device", choices=["device", "host"], help="where slice data is parsed: kernel Kp (default) or the host cores")
    ap.add_argument("--shard", default="streams", choices=["streams", "gop"], help="gop: strong scaling over the IDR-bounded GOP segments of a fixed set of 4K streams (SURVEY 8e; implies --workload 4k_gop)")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--e2e-only", action="store_true", help="host-to-host leg only (experiments)")
    ap.add_argument("--skip-e2e", action="store_true", help="kernel-tuning aid: skip the host-to-host leg (the line then carries no valid e2e and is not a bench result)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.shard == "gop":
        args.workload = "4k_gop"
    _WL["name"] = args.workload
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        # the reference arm loads nothing of the product: only the stream writer (input generator) and oracle/_ref/refdec
        if rank == 0:
            from broadway_b200 import build as b
            b.build_writer()
            b.build_oracle()
        reference_arm(args, rank, world)
        return
    import __graft_entry__
    if local_rank == 0:
        __graft_entry__.build()
    if args.warmup < 3:
        log("note: fewer than 3 warm-up steps requested")
    if args.shard == "gop":
        gop_arm(args, rank, local_rank, world)
    else:
        own_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
