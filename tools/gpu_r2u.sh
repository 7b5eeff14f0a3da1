# round 2, GPU call U: split of the SMs between Kp and K1..K4, and the host share, under the free-running pipeline
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
for x in 92 100 106; do echo "kp_sms $x"; H264B200_KP_SMS=$x timeout 600 $E 2>> gpurun_out/r2u_e2e.log | tee gpurun_out/r2u_e2e_x$x.json; done
for h in 32 56; do echo "kp_sms 100 host $h"; H264B200_KP_SMS=100 H264B200_HOST_STREAMS=$h timeout 600 $E 2>> gpurun_out/r2u_e2e.log | tee gpurun_out/r2u_e2e_x100_h$h.json; done
echo "kp_sms 84 host 56"; H264B200_KP_SMS=84 H264B200_HOST_STREAMS=56 H264B200_TIMELINE=gpurun_out/r2u_timeline_x84_h56.csv timeout 600 $E 2>> gpurun_out/r2u_e2e.log | tee gpurun_out/r2u_e2e_x84_h56.json
