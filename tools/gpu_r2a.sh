# round 2, first GPU call of this session: GPU tests, default bench (device parse), host-parse e2e for comparison, reference arm, ncu of Kp
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2a_gputests.log; tail -5 gpurun_out/r2a_gputests.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.log; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.log; cat gpurun_out/r2a_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.log; cat gpurun_out/r2a_ref.json
CMD="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --streams 256"
timeout 600 $CMD > gpurun_out/ncu_plain_kp.json 2> gpurun_out/ncu_plain_kp.log &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kp_parse -s 2 -c 1 -f -o gpurun_out/r2a_kp_256 $CMD > gpurun_out/ncu_kp_256.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_kp_256.log
