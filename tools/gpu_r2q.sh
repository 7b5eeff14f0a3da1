# round 2, GPU call Q: scratch of the launches sized once (no cudaFree under load): e2e + timeline; K2 what-if variants (window loads free / interpolation free)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2q_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2q_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2q_timeline.csv timeout 600 $E 2>> gpurun_out/r2q_e2e.log | tee gpurun_out/r2q_e2e_default.json
H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2q_e2e.log | tee gpurun_out/r2q_e2e_x120.json
timeout 600 python bench.py --e2e-only --no-check --steps 6 --warmup 3 2>> gpurun_out/r2q_e2e.log | tee gpurun_out/r2q_e2e_steps6.json
grep scheduling gpurun_out/r2q_e2e.log | tail -2
B="python bench.py --parse host --skip-e2e --no-check --no-cpu-baseline --steps 3 --warmup 1 --frames 16"
for v in base k2nowin k2nomath; do
  if [ $v = base ]; then timeout 600 $B > gpurun_out/r2q_k2_$v.json 2>> gpurun_out/r2q_k2.log; else H264B200_LIB=build/variants/libh264b200_$v.so timeout 600 $B > gpurun_out/r2q_k2_$v.json 2>> gpurun_out/r2q_k2.log; fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2q_k2_$v.json")); k=d["roofline"]["kernels"]
    print("$v:", {n: round(k[n]["ms_per_launch"],3) for n in k})
except Exception as e: print("$v failed", e)
PY
done
