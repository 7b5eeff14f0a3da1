# round 2, GPU call E: pipelined consume (two rounds in flight), stage parity tests, GOP-sharded arm at N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gputests.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2e_gputests.log; tail -5 gpurun_out/r2e_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
rm -f gpurun_out/r2e_timeline_*.csv
echo "w16 c4"; H264B200_TIMELINE=gpurun_out/r2e_timeline_w16c4.csv timeout 600 $E 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_w16c4.json
echo "w32 c4"; H264B200_WINDOW=32 H264B200_KP_CHUNK=4 H264B200_TIMELINE=gpurun_out/r2e_timeline_w32c4.csv timeout 600 $E 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_w32c4.json
echo "w32 c8"; H264B200_WINDOW=32 H264B200_KP_CHUNK=8 timeout 600 $E 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_w32c8.json
echo "w24 c2"; H264B200_WINDOW=24 H264B200_KP_CHUNK=2 timeout 600 $E 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_w24c2.json
echo "8 threads w16 c4"; timeout 600 $E --threads 8 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_t8.json
echo "4 threads w16 c4"; timeout 600 $E --threads 4 2>> gpurun_out/r2e_e2e.log | tee gpurun_out/r2e_e2e_t4.json
timeout 900 python bench.py --shard gop --steps 2 --warmup 1 > gpurun_out/r2e_gop_n1.json 2> gpurun_out/r2e_gop_n1.log; echo "gop exit $?"; tail -2 gpurun_out/r2e_gop_n1.log; cat gpurun_out/r2e_gop_n1.json | cut -c1-600
