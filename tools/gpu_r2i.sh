# round 2, GPU call I: exclusive Kp launches (kp_parse<32,1>: one CTA per SM, 112 SMs) against shared mode; window/chunk from the engine
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2i_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2i_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
rm -f gpurun_out/r2i_timeline*.csv
echo "exclusive 112 SMs (default: chunk 14, window 32)"; H264B200_TIMELINE=gpurun_out/r2i_timeline_x112.csv timeout 600 $E 2>> gpurun_out/r2i_e2e.log | tee gpurun_out/r2i_e2e_x112.json
echo "exclusive 104 SMs"; H264B200_KP_SMS=104 timeout 600 $E 2>> gpurun_out/r2i_e2e.log | tee gpurun_out/r2i_e2e_x104.json
echo "exclusive 120 SMs"; H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2i_e2e.log | tee gpurun_out/r2i_e2e_x120.json
echo "shared (kp_sms 0, w32 c8)"; H264B200_KP_SMS=0 H264B200_WINDOW=32 H264B200_KP_CHUNK=8 timeout 600 $E 2>> gpurun_out/r2i_e2e.log | tee gpurun_out/r2i_e2e_shared.json
tail -3 gpurun_out/r2i_e2e.log
