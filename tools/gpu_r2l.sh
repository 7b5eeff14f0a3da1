# round 2, GPU call L: the host share of h264b200DecodeStreams (streams parsed by the idle worker threads instead of Kp): tests + sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2l_gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2l_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
for h in default 0 40 64 80; do
  echo "host streams $h"
  if [ $h = default ]; then timeout 600 $E 2>> gpurun_out/r2l_e2e.log | tee gpurun_out/r2l_e2e_h$h.json
  else H264B200_HOST_STREAMS=$h timeout 600 $E 2>> gpurun_out/r2l_e2e.log | tee gpurun_out/r2l_e2e_h$h.json; fi
done
H264B200_TIMELINE=gpurun_out/r2l_timeline_default.csv timeout 600 $E 2>> gpurun_out/r2l_e2e.log > /dev/null
tail -3 gpurun_out/r2l_e2e.log
