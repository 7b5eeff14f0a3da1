# round 2, GPU call B: tests with the new cases, default bench with the chunked Kp pipeline (64 pictures per stream), wavefront
# acquire A/B, Kp throughput against pictures in flight, PCIe / NUMA probe
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | head -30; numactl -H 2>/dev/null; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; free -g) > gpurun_out/r2b_box.txt 2>&1
timeout 300 build/pcie_probe > gpurun_out/r2b_pcie.json 2> gpurun_out/r2b_pcie.log; echo "pcie exit $?"; cat gpurun_out/r2b_pcie.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputests.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2b_gputests.log; tail -5 gpurun_out/r2b_gputests.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.log; echo "bench exit $?"; tail -3 gpurun_out/r2b_bench.log; python tools/show_bench.py gpurun_out/r2b_bench.json
B="python bench.py --skip-e2e --no-check --no-cpu-baseline --steps 2 --warmup 1 --frames 16"
for v in wfacq0 wfacq2; do H264B200_LIB=build/variants/libh264b200_$v.so timeout 600 $B > gpurun_out/r2b_$v.json 2> gpurun_out/r2b_$v.log; echo $v; python tools/show_bench.py gpurun_out/r2b_$v.json; done
for S in 64 128 512; do timeout 600 $B --streams $S > gpurun_out/r2b_kp_$S.json 2> gpurun_out/r2b_kp_$S.log; echo streams $S; python tools/show_bench.py gpurun_out/r2b_kp_$S.json; done
timeout 600 python bench.py --e2e-only --no-check --steps 3 --warmup 1 --frames 16 > gpurun_out/r2b_e2e_f16.json 2>> gpurun_out/r2b_bench.log; cat gpurun_out/r2b_e2e_f16.json
H264B200_WINDOW=32 timeout 600 python bench.py --e2e-only --no-check --steps 3 --warmup 1 > gpurun_out/r2b_e2e_w32.json 2>> gpurun_out/r2b_bench.log; cat gpurun_out/r2b_e2e_w32.json
H264B200_WINDOW=8 timeout 600 python bench.py --e2e-only --no-check --steps 3 --warmup 1 > gpurun_out/r2b_e2e_w8.json 2>> gpurun_out/r2b_bench.log; cat gpurun_out/r2b_e2e_w8.json
