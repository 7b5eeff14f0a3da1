# round 2, GPU call D: where the e2e pipeline loses time: device timeline of the default run, window / chunk sweep (K steps streamed in one call)
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
rm -f gpurun_out/r2d_timeline_w16c4.csv
echo "w16 c4"; H264B200_TIMELINE=gpurun_out/r2d_timeline_w16c4.csv timeout 600 $E 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_w16c4.json
echo "w32 c4"; H264B200_WINDOW=32 H264B200_KP_CHUNK=4 H264B200_TIMELINE=gpurun_out/r2d_timeline_w32c4.csv timeout 600 $E 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_w32c4.json
echo "w32 c2"; H264B200_WINDOW=32 H264B200_KP_CHUNK=2 timeout 600 $E 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_w32c2.json
echo "w24 c4"; H264B200_WINDOW=24 H264B200_KP_CHUNK=4 timeout 600 $E 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_w24c4.json
echo "w32 c8"; H264B200_WINDOW=32 H264B200_KP_CHUNK=8 timeout 600 $E 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_w32c8.json
echo "host parse"; timeout 600 $E --parse host 2>> gpurun_out/r2d_e2e.log | tee gpurun_out/r2d_e2e_host.json
tail -5 gpurun_out/r2d_e2e.log
