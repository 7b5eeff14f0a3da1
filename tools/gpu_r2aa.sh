# round 2, GPU call AA: Kp launches in multiples of 32 pictures, replay with Kp's SM share: default bench (short)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_parse_gpu.py -x -q > gpurun_out/r2aa_gputests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2aa_gputests.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2aa_bench.err | tee gpurun_out/r2aa_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],{k:round(v['ms_per_launch'],2) for k,v in d['roofline']['kernels'].items()})"
