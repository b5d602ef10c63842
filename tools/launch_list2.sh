# usage: bash tools/launch_list2.sh <tag> "ENV=val ..." [skip] [count] -- ncu durations (no cache flush between kernels) of a window of launches
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; envs=$2; skip=${3:-300}; cnt=${4:-96}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8"
env $envs ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s $skip -c $cnt --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
python - <<PY
import csv, io, collections
raw=[l for l in open("gpurun_out/${tag}_launches.csv") if not l.startswith("==")]
rows=list(csv.DictReader(io.StringIO("".join(raw))))
agg=collections.OrderedDict()
for r in rows:
    k=r["Kernel Name"].split("(")[0][:60]
    a=agg.setdefault(k,[]); a.append(float(r["Metric Value"].replace(",",""))/1e3)
for k,v in agg.items(): print(k, len(v), "mean %.1f us"%(sum(v)/len(v)), "min %.1f max %.1f"%(min(v),max(v)))
PY
