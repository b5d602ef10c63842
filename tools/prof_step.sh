# usage: bash tools/prof_step.sh <tag> [kernel-regex]   -- one ncu --set full capture of the step kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-x}; kre=${2:-step_kernel_async}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o gpurun_out/$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
