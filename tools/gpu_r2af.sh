# round 2, GPU call AF: second reconstruction lane (H264B200_SPLIT_ROUNDS=1, off by default): default path re-verified, then the experiment
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2af_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2af_gputests.log
E="python bench.py --e2e-only --no-check --steps 3 --warmup 1"
timeout 120 $E 2>> gpurun_out/r2af_e2e.log | tee gpurun_out/r2af_e2e_default.json
H264B200_SPLIT_ROUNDS=1 timeout 120 $E 2>> gpurun_out/r2af_e2e.log | tee gpurun_out/r2af_e2e_split.json
H264B200_SPLIT_ROUNDS=1 timeout 120 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2af_gputests_split.log 2>&1; echo "pytest(split) exit $?"; tail -2 gpurun_out/r2af_gputests_split.log
