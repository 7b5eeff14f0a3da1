# round 2, GPU call P: spare frame slot, copy-out issued outside the engine mutex, launch publication fix: all GPU tests, e2e A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2p_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2p_timeline.csv timeout 600 $E 2>> gpurun_out/r2p_e2e.log | tee gpurun_out/r2p_e2e_default.json
H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2p_e2e.log | tee gpurun_out/r2p_e2e_x120.json
timeout 600 python bench.py --e2e-only --no-check --steps 6 --warmup 3 2>> gpurun_out/r2p_e2e.log | tee gpurun_out/r2p_e2e_steps6.json
tail -3 gpurun_out/r2p_e2e.log
