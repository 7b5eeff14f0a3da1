# round 2, GPU call AD: does the round size (streams per GPU) move the end-to-end rate?
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
for s in 384 512; do echo "streams $s"; H264B200_TIMELINE=gpurun_out/r2ad_timeline_s$s.csv timeout 600 $E --streams $s 2>> gpurun_out/r2ad_e2e.log | tee gpurun_out/r2ad_e2e_s$s.json; done
grep 'h264b200 ' gpurun_out/r2ad_e2e.log | tail -4
