# round 2, GPU call AE: the final commit once more: all GPU tests, smoke, a short default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2ae_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ae_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2ae_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2ae_bench.err | tee gpurun_out/r2ae_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],d['parity'])"
