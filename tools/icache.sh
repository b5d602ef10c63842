# usage: bash tools/icache.sh <tag> [env settings...]  -- instruction-delivery counters of one step-kernel launch (ncu, metrics only)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; shift
env "$@" timeout 600 ncu --metrics gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
  --clock-control none -k regex:step_kernel_async -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep -E "gcc__|sm__|gpu__|smsp__" | awk '{print $1, $NF}' | tr '\n' ' ' > gpurun_out/${tag}_icache.log
echo >> gpurun_out/${tag}_icache.log; echo "$tag $*"; cat gpurun_out/${tag}_icache.log
