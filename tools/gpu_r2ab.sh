# round 2, GPU call AB (2 GPUs): the driver's multi-GPU launch line at N=2, both arms: does end-to-end scale now?
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2ab_topo.txt 2>&1; nproc >> gpurun_out/r2ab_topo.txt; lscpu | grep -i 'numa\|socket\|model name' >> gpurun_out/r2ab_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2> gpurun_out/r2ab_ref.err | tee gpurun_out/r2ab_bench_reference_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/r2ab_bench.err | tee gpurun_out/r2ab_bench_2gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 value',d['value'],'e2e',d['e2e']['value'],d['config'].get('numa'),d['config']['parser_threads_per_gpu'])"
tail -4 gpurun_out/r2ab_bench.err
