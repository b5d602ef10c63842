# usage: bash tools/bench_envs.sh <tag>  -- bench.py --env <id> for all seven registered ids
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-x}
: > gpurun_out/${tag}_bench_envs.jsonl
for id in GripperTouch-v0 BlocksTouch-v0 BlocksTouchCurriculum-v0 BlocksTouchChoose-v0 BlocksTouchChooseCurriculum-v0 BlocksTouchVariation-v0 ToppleTower-v0; do
  timeout 300 python bench.py --env $id --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>/dev/null | grep '^{' >> gpurun_out/${tag}_bench_envs.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_envs.jsonl"):
    d=json.loads(l); es=d["episode_stats"]; print(d["config"]["env_id"], "%.3g"%d["value"], "%.2f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "full %.3f"%(es["worker_steps"]/es["steps"]))
PY
