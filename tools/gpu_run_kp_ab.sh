timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gputests_2.log; tail -12 gpurun_out/r2_gputests_2.log
B="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1"
for cfg in "base 256" "base 64" "base 512"; do set -- $cfg; $B --streams $2 > gpurun_out/r2_kp_$1_$2.json 2> gpurun_out/r2_kp_$1_$2.log; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_kp_$1_$2.json")); k=d["roofline"]["kernels"]["kp_parse"]; print("$1 streams $2: kp ms/launch %.1f launches %d  value %.0f" % (k["ms_per_launch"], k["launches"], d["value"]))
except Exception as e: print("$1 $2 failed", e)
PY
done
for v in kpb2 kpw4b8; do H264B200_LIB=build/variants/libh264b200_$v.so $B --streams 256 > gpurun_out/r2_kp_${v}_256.json 2> gpurun_out/r2_kp_${v}_256.log; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_kp_${v}_256.json")); k=d["roofline"]["kernels"]["kp_parse"]; print("$v streams 256: kp ms/launch %.1f launches %d  value %.0f" % (k["ms_per_launch"], k["launches"], d["value"]))
except Exception as e: print("$v failed", e)
PY
done
