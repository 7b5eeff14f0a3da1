# round 2, GPU call M: timeline of the all-device e2e run (what holds it below the replay's rate), ncu --set full of kernel Kp
mkdir -p gpurun_out
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_HOST_STREAMS=0 H264B200_TIMELINE=gpurun_out/r2m_timeline_h0.csv timeout 600 $E 2>> gpurun_out/r2m_e2e.log | tee gpurun_out/r2m_e2e_h0.json
H264B200_HOST_STREAMS=0 H264B200_KP_SMS=128 timeout 600 $E 2>> gpurun_out/r2m_e2e.log | tee gpurun_out/r2m_e2e_h0_x128.json
H264B200_HOST_STREAMS=0 H264B200_WINDOW=28 timeout 600 $E 2>> gpurun_out/r2m_e2e.log | tee gpurun_out/r2m_e2e_h0_w28.json
CMD="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --frames 16"
H264B200_KP_SMS=0 H264B200_WINDOW=16 H264B200_KP_CHUNK=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kp_parse -s 1 -c 1 -f -o gpurun_out/r2m_kp $CMD > gpurun_out/r2m_ncu_kp.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2m_ncu_kp.log; ls -la gpurun_out/*.ncu-rep
