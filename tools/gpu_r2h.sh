# round 2, GPU call H: stream priorities (reconstruction high, Kp low) on / off, branch-free residual loop head; then the full default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2h_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2h_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
rm -f gpurun_out/r2h_timeline*.csv
echo "prio w16 c4"; H264B200_TIMELINE=gpurun_out/r2h_timeline_prio_w16c4.csv timeout 600 $E 2>> gpurun_out/r2h_e2e.log | tee gpurun_out/r2h_e2e_prio_w16c4.json
echo "noprio w16 c4"; H264B200_PRIO=0 timeout 600 $E 2>> gpurun_out/r2h_e2e.log | tee gpurun_out/r2h_e2e_noprio_w16c4.json
echo "prio w32 c8"; H264B200_WINDOW=32 H264B200_KP_CHUNK=8 timeout 600 $E 2>> gpurun_out/r2h_e2e.log | tee gpurun_out/r2h_e2e_prio_w32c8.json
echo "prio w32 c16"; H264B200_WINDOW=32 H264B200_KP_CHUNK=16 timeout 600 $E 2>> gpurun_out/r2h_e2e.log | tee gpurun_out/r2h_e2e_prio_w32c16.json
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.log; echo "bench exit $?"; tail -2 gpurun_out/r2h_bench.log; python tools/show_bench.py gpurun_out/r2h_bench.json
