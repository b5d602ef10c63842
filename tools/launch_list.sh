# usage: bash tools/launch_list.sh <tag> "ENV=val ..."  -- ncu launch list (durations) of a short bench run under the given environment
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; shift
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8"
env $1 $CMD > gpurun_out/${tag}_plain.log 2>&1 && \
env $1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
python - <<PY
import csv, io, collections
raw=[l for l in open("gpurun_out/${tag}_launches.csv") if not l.startswith("==")]
rows=list(csv.DictReader(io.StringIO("".join(raw))))
agg=collections.OrderedDict()
for r in rows:
    k=(r["Kernel Name"].split("(")[0][:60], r["Metric Name"])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"].replace(",",""))
for (k,m),(n,v) in agg.items(): print(k, m, n, "total %.4g"%v, "mean %.4g"%(v/n))
PY
