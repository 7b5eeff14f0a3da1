# round 2, GPU call V: workers no longer query CUDA events (the scheduling thread publishes what it sees): tests, e2e, host share
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2v_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2v_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2v_timeline.csv timeout 600 $E 2>> gpurun_out/r2v_e2e.log | tee gpurun_out/r2v_e2e_default.json
H264B200_HOST_STREAMS=32 timeout 600 $E 2>> gpurun_out/r2v_e2e.log | tee gpurun_out/r2v_e2e_h32.json
H264B200_HOST_STREAMS=56 timeout 600 $E 2>> gpurun_out/r2v_e2e.log | tee gpurun_out/r2v_e2e_h56.json
timeout 600 python bench.py --e2e-only --no-check --steps 3 --warmup 1 --threads 4 2>> gpurun_out/r2v_e2e.log | tee gpurun_out/r2v_e2e_t4.json
grep scheduling gpurun_out/r2v_e2e.log | tail -2
