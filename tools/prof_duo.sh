cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for mode in duo async; do echo $mode; BP_STEP_KERNEL=$mode python tools/stats_probe.py; done > gpurun_out/stats_probe.log 2>&1
cat gpurun_out/stats_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_duo -s 3 -c 1 -o gpurun_out/duo1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/duo1_ncu.log 2>&1
tail -3 gpurun_out/duo1_ncu.log
