# usage: bash tools/prof_split.sh <tag>  -- per-kernel durations and key counters of the split step kernels (ncu, a few launches)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1
BP_STEP_KERNEL=split timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
  --clock-control none -k regex:split_ -s 600 -c 9 --csv --log-file gpurun_out/${tag}_split.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 > gpurun_out/${tag}_split.log 2>&1
python - <<PY
import csv
rows=list(csv.DictReader(open("gpurun_out/${tag}_split.csv")))
from collections import defaultdict
d=defaultdict(dict)
for r in rows:
    d[(r["ID"], r["Kernel Name"][:40])][r["Metric Name"]]=r["Metric Value"]
for k,v in d.items():
    print(k[1], {m.split("__")[-1][:34]:x for m,x in v.items()})
PY
