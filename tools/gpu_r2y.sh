# round 2, GPU call Y: host share with workers of its own (3 scanning workers keep the device-parsed streams)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2y_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2y_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
for h in auto 32 64; do echo "host $h"; H264B200_HOST_STREAMS=$h H264B200_TIMELINE=gpurun_out/r2y_timeline_h$h.csv timeout 600 $E 2>> gpurun_out/r2y_e2e.log | tee gpurun_out/r2y_e2e_h$h.json; done
echo "host auto kp_sms 96"; H264B200_KP_SMS=96 H264B200_HOST_STREAMS=auto timeout 600 $E 2>> gpurun_out/r2y_e2e.log | tee gpurun_out/r2y_e2e_hauto_x96.json
grep 'h264b200 ' gpurun_out/r2y_e2e.log | tail -8
