# round 2, GPU call G: Kp after pinning the shared-memory bases and branching around the refill; e2e default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2g_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2g_gputests.log
B="python bench.py --skip-e2e --no-check --no-cpu-baseline --steps 2 --warmup 1 --frames 16"
H264B200_WINDOW=16 H264B200_KP_CHUNK=16 timeout 600 $B > gpurun_out/r2g_kp_base.json 2> gpurun_out/r2g_kp_base.log; echo base; python tools/show_bench.py gpurun_out/r2g_kp_base.json
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
echo "e2e w16 c4"; timeout 600 $E 2>> gpurun_out/r2g_e2e.log | tee gpurun_out/r2g_e2e_w16c4.json
echo "e2e w32 c8"; H264B200_WINDOW=32 H264B200_KP_CHUNK=8 timeout 600 $E 2>> gpurun_out/r2g_e2e.log | tee gpurun_out/r2g_e2e_w32c8.json
