"""Summarise an engine timeline (H264B200_TIMELINE=file.csv): what the GPU did in the steady-state third of the run.
usage: python tools/timeline_summary.py gpurun_out/timeline.csv [kp_sms]"""
import collections, csv, sys
path = sys.argv[1]; kp_sms = int(sys.argv[2]) if len(sys.argv) > 2 else 112
by = collections.defaultdict(list)
for r in csv.DictReader(open(path)):
    if r["kind"] == "kind":
        by = collections.defaultdict(list)          # a later run appended to the same file: keep the last one
        continue
    by[r["kind"]].append((int(r["pictures"]), float(r["host_launch_ms"]), float(r["gpu_start_ms"]), float(r["gpu_end_ms"])))
R = by["round"]
if not R:
    sys.exit("no rounds")
mid = R[len(R) // 3: 2 * len(R) // 3]
t0, t1 = mid[0][2], mid[-1][3]
span = t1 - t0
print("rounds %d, Kp launches %d; steady-state window: %d rounds in %.0f ms = %.2f ms per round = %.0f pictures/s" %
      (len(R), len(by["kp"]), len(mid), span, span / len(mid), 1000.0 * sum(p for p, _, _, _ in mid) / span))
print("  round kernels: mean %.2f ms, busy %.0f%% of the window; gap between rounds mean %.2f ms" %
      (sum(e - s for _, _, s, e in mid) / len(mid), 100.0 * sum(e - s for _, _, s, e in mid) / span,
       sum(mid[i + 1][2] - mid[i][3] for i in range(len(mid) - 1)) / max(1, len(mid) - 1)))
print("  round: host launch -> GPU start mean %.2f ms" % (sum(s - h for _, h, s, _ in mid) / len(mid)))
K = [k for k in by["kp"] if k[3] > t0 and k[2] < t1]
if K:
    busy = sum(min(k[3], t1) - max(k[2], t0) for k in K)
    ctas = sum((min(k[3], t1) - max(k[2], t0)) * min(kp_sms, (k[0] + 31) // 32) for k in K) / span
    print("  Kp: %d launches overlap the window, mean %.0f pictures, mean duration %.1f ms, %.2f in flight on average = %.0f SMs of %d; host launch -> GPU start mean %.1f ms" %
          (len(K), sum(k[0] for k in K) / len(K), sum(k[3] - k[2] for k in K) / len(K), busy / span, ctas, kp_sms, sum(k[2] - k[1] for k in K) / len(K)))
    print("  Kp pictures finished in the window: %.0f per s" % (1000.0 * sum(k[0] for k in K if t0 <= k[3] <= t1) / span))
D = [d for d in by["d2h"] if d[2] >= t0 and d[3] <= t1]
if D:
    print("  copy-out: mean %.2f ms per round, busy %.0f%% of the window" % (sum(d[3] - d[2] for d in D) / len(D), 100.0 * sum(d[3] - d[2] for d in D) / span))
