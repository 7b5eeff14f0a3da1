# round 2, GPU call T: a slot's copy-out event is only waited for / polled while its scratch set still serves that round: tests, e2e, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2t_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
H264B200_TIMELINE=gpurun_out/r2t_timeline.csv timeout 600 $E 2>> gpurun_out/r2t_e2e.log | tee gpurun_out/r2t_e2e_default.json
H264B200_KP_SMS=120 timeout 600 $E 2>> gpurun_out/r2t_e2e.log | tee gpurun_out/r2t_e2e_x120.json
timeout 600 python bench.py --e2e-only --no-check --steps 6 --warmup 3 2>> gpurun_out/r2t_e2e.log | tee gpurun_out/r2t_e2e_steps6.json
grep scheduling gpurun_out/r2t_e2e.log | tail -2
