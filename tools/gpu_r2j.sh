# round 2, GPU call J: exclusive Kp launches as many small launches within an SM budget; sweeps of pictures per launch and of the SM share
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2j_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2j_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
rm -f gpurun_out/r2j_timeline*.csv
echo "x112 c2 (default)"; H264B200_TIMELINE=gpurun_out/r2j_timeline_x112c2.csv timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x112c2.json
echo "x112 c1"; H264B200_KP_CHUNK=1 timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x112c1.json
echo "x112 c4 w24"; H264B200_KP_CHUNK=4 H264B200_WINDOW=24 timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x112c4.json
echo "x104 c2"; H264B200_KP_SMS=104 timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x104c2.json
echo "x120 c2"; H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x120c2.json
echo "x96 c2"; H264B200_KP_SMS=96 timeout 600 $E 2>> gpurun_out/r2j_e2e.log | tee gpurun_out/r2j_e2e_x96c2.json
tail -3 gpurun_out/r2j_e2e.log
