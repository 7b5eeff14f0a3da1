# round 2, GPU call K: state of the tree after the container was re-created: all GPU tests, smoke, default bench (both arms), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2k_box.txt; nproc >> gpurun_out/r2k_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_gputests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2k_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2k_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2k_ref.err | tee gpurun_out/r2k_bench_reference.json
timeout 900 python bench.py 2> gpurun_out/r2k_bench.err | tee gpurun_out/r2k_bench_default.json
tail -5 gpurun_out/r2k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_resident.csv python bench.py --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --frames 16 > gpurun_out/r2k_ncu.log 2>&1; echo "ncu exit $?"
