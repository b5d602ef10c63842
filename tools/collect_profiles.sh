# usage: bash tools/collect_profiles.sh <tag>  -- the round's bench lines, ncu launch list and one ncu --set full capture
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-r2}
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_1gpu.err
python bench_her.py > gpurun_out/${tag}_bench_her.jsonl 2>> gpurun_out/${tag}_bench_1gpu.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_async -s 3 -c 1 -o gpurun_out/${tag}_step_kernel -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log; cat gpurun_out/${tag}_bench_1gpu.json | cut -c1-400; tail -3 gpurun_out/${tag}_bench_1gpu.err
: > gpurun_out/${tag}_bench_envs.jsonl
for id in GripperTouch-v0 BlocksTouch-v0 BlocksTouchCurriculum-v0 BlocksTouchChoose-v0 BlocksTouchChooseCurriculum-v0 BlocksTouchVariation-v0 ToppleTower-v0; do
  timeout 300 python bench.py --env $id --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>/dev/null | grep '^{' >> gpurun_out/${tag}_bench_envs.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_envs.jsonl"):
    d=json.loads(l); print(d["config"]["env_id"], "%.3g"%d["value"], "%.4f"%d["roofline"]["frac"])
PY
