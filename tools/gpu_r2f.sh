# round 2, GPU call F: ncu --set full of Kp after the rewrite (one launch of 4096 pictures, comparable with r2a), e2e serialised variants
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
echo "w32 c16 serialised"; H264B200_WINDOW=32 H264B200_KP_CHUNK=16 H264B200_KP_ON_COMP=1 timeout 600 $E 2>> gpurun_out/r2f_e2e.log | tee gpurun_out/r2f_e2e_w32c16s.json
echo "w32 c8 serialised"; H264B200_WINDOW=32 H264B200_KP_CHUNK=8 H264B200_KP_ON_COMP=1 timeout 600 $E 2>> gpurun_out/r2f_e2e.log | tee gpurun_out/r2f_e2e_w32c8s.json
echo "w48 c16 overlapped"; H264B200_WINDOW=48 H264B200_KP_CHUNK=16 timeout 600 $E 2>> gpurun_out/r2f_e2e.log | tee gpurun_out/r2f_e2e_w48c16o.json
export H264B200_WINDOW=16 H264B200_KP_CHUNK=16
CMD="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --streams 256 --frames 16"
timeout 600 $CMD > gpurun_out/ncu_plain_kp.json 2> gpurun_out/ncu_plain_kp.log &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kp_parse -s 2 -c 1 -f -o gpurun_out/r2f_kp_256 $CMD > gpurun_out/ncu_kp_256.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_kp_256.log
