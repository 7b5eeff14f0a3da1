# round 2, GPU call AC (2 GPUs): BASELINE configs[4], IDR-bounded GOP segments of 4K streams sharded over the GPUs (strong scaling), N=1 then N=2
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --shard gop --gpus 1 --steps 3 --warmup 1 2> gpurun_out/r2ac_gop1.err | tee gpurun_out/r2ac_gop_1gpu.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --shard gop --gpus 2 --steps 3 --warmup 1 2> gpurun_out/r2ac_gop2.err | tee gpurun_out/r2ac_gop_2gpu.json | cut -c1-400
tail -3 gpurun_out/r2ac_gop1.err gpurun_out/r2ac_gop2.err
