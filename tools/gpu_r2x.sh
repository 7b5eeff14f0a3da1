# round 2, GPU call X: outputs collected every 0.5 ms while scanning: tests, e2e at 16 / 4 threads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2x_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2x_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2x_timeline.csv timeout 600 $E 2>> gpurun_out/r2x_e2e.log | tee gpurun_out/r2x_e2e_default.json
timeout 600 $E --threads 4 2>> gpurun_out/r2x_e2e.log | tee gpurun_out/r2x_e2e_t4.json
H264B200_HOST_STREAMS=32 timeout 600 $E 2>> gpurun_out/r2x_e2e.log | tee gpurun_out/r2x_e2e_h32.json
grep 'h264b200 ' gpurun_out/r2x_e2e.log | tail -3
