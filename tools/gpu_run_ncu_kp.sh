# ncu --set full of kernel Kp on the resident replay of the bench (argument: streams per GPU)
S=${1:-256}
CMD="python bench.py --parse device --skip-e2e --no-check --no-cpu-baseline --steps 1 --warmup 1 --streams $S"
$CMD > gpurun_out/ncu_plain_kp.json 2> gpurun_out/ncu_plain_kp.log &&
ncu --set full --clock-control none --import-source on -k regex:kp_parse -s 2 -c 1 -f -o gpurun_out/r2_kp_$S $CMD > gpurun_out/ncu_kp_$S.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_kp_$S.log
