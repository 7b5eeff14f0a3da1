# round 2, GPU call O: free-running pipeline, second version (cheap polls, idle-only relaxation, event-reuse fix): tests, e2e, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2o_gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2o_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2o_timeline.csv timeout 600 $E 2>> gpurun_out/r2o_e2e.log | tee gpurun_out/r2o_e2e_default.json
H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2o_e2e.log | tee gpurun_out/r2o_e2e_x120.json
H264B200_KP_CHUNK=1 timeout 600 $E 2>> gpurun_out/r2o_e2e.log | tee gpurun_out/r2o_e2e_c1.json
H264B200_HOST_STREAMS=32 timeout 600 $E 2>> gpurun_out/r2o_e2e.log | tee gpurun_out/r2o_e2e_h32.json
tail -3 gpurun_out/r2o_e2e.log
