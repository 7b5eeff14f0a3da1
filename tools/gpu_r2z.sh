# round 2, GPU call Z: final state: all GPU tests, smoke, both bench arms with the driver's command line, launch list of the resident region
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2z_box.txt; nproc >> gpurun_out/r2z_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2z_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2z_ref.err | tee gpurun_out/r2z_bench_reference.json
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2z_bench.err | tee gpurun_out/r2z_bench_default.json
tail -3 gpurun_out/r2z_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_resident.csv python bench.py --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --frames 16 > gpurun_out/r2z_ncu.log 2>&1; echo "ncu exit $?"
