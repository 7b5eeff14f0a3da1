# quick Kp timing on the resident replay (no e2e, no checks): prints ms per Kp launch for the given stream counts
B="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1"
for S in ${@:-256}; do
  $B --streams $S > gpurun_out/kpq_$S.json 2> gpurun_out/kpq_$S.log || tail -5 gpurun_out/kpq_$S.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/kpq_$S.json")); k=d["roofline"]["kernels"]["kp_parse"]
    print("streams $S: kp %.1f ms/launch x%d = %.0f pictures/s ; value %.0f" % (k["ms_per_launch"], k["launches"], $S*16/k["ms_per_launch"]*1000, d["value"]))
except Exception as e: print("$S failed", e)
PY
done
