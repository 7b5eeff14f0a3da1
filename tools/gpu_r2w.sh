# round 2, GPU call W: why rounds are not full (diagnostic counters), threads 4 / 8 / 16
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2w_timeline.csv timeout 600 $E 2>> gpurun_out/r2w_e2e.log | tee gpurun_out/r2w_e2e_default.json
H264B200_TIMELINE=gpurun_out/r2w_timeline_t6.csv timeout 600 $E --threads 6 2>> gpurun_out/r2w_e2e.log | tee gpurun_out/r2w_e2e_t6.json
grep 'h264b200 ' gpurun_out/r2w_e2e.log | tail -6
