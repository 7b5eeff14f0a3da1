# round 2, GPU call AG: the end-to-end rate of the final default path over a longer run
mkdir -p gpurun_out
timeout 170 python bench.py --e2e-only --no-check --steps 14 --warmup 3 2>> gpurun_out/r2ag_e2e.log | tee gpurun_out/r2ag_e2e_steps14.json
