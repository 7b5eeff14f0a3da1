# round 2, GPU call C: Kp timing after the bit reader / residual loop rewrite (one launch of 4096 pictures), register variants,
# scheduling policies of the e2e pipeline
mkdir -p gpurun_out
B="python bench.py --skip-e2e --no-check --no-cpu-baseline --steps 2 --warmup 1 --frames 16"
export H264B200_WINDOW=16 H264B200_KP_CHUNK=16
timeout 600 $B > gpurun_out/r2c_kp_base.json 2> gpurun_out/r2c_kp_base.log; echo base; python tools/show_bench.py gpurun_out/r2c_kp_base.json
for v in kpr80 kpr48; do H264B200_LIB=build/variants/libh264b200_$v.so timeout 600 $B > gpurun_out/r2c_$v.json 2> gpurun_out/r2c_$v.log; echo $v; python tools/show_bench.py gpurun_out/r2c_$v.json; done
unset H264B200_WINDOW H264B200_KP_CHUNK
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
echo "e2e default (window 16, chunk 4, overlapped)"; timeout 600 $E 2>> gpurun_out/r2c_e2e.log | tee gpurun_out/r2c_e2e_default.json
echo "e2e window 32 chunk 16 serialised"; H264B200_WINDOW=32 H264B200_KP_CHUNK=16 H264B200_KP_ON_COMP=1 timeout 600 $E 2>> gpurun_out/r2c_e2e.log | tee gpurun_out/r2c_e2e_w32c16s.json
echo "e2e window 32 chunk 16 overlapped"; H264B200_WINDOW=32 H264B200_KP_CHUNK=16 timeout 600 $E 2>> gpurun_out/r2c_e2e.log | tee gpurun_out/r2c_e2e_w32c16o.json
echo "e2e window 16 chunk 8 serialised"; H264B200_WINDOW=16 H264B200_KP_CHUNK=8 H264B200_KP_ON_COMP=1 timeout 600 $E 2>> gpurun_out/r2c_e2e.log | tee gpurun_out/r2c_e2e_w16c8s.json
echo "e2e window 16 chunk 4 serialised"; H264B200_KP_ON_COMP=1 timeout 600 $E 2>> gpurun_out/r2c_e2e.log | tee gpurun_out/r2c_e2e_w16c4s.json
