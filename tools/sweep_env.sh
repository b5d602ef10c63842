# usage: bash tools/sweep_env.sh <tag> "VAR=val VAR2=val" "VAR=val" ...   -- one short bench line per environment setting
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; shift
: > gpurun_out/${tag}_sweep.log
for setting in "$@"; do
  env $setting timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']; n=(d['steps']+d['warmup'])*8192
print('$setting', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'iters/slab %.1f passes/slab %.1f fill %.1f'%(e['sched_iterations']/n, e['sched_passes']/n, e['worker_steps']/max(e['sched_passes'],1)))" >> gpurun_out/${tag}_sweep.log
done
cat gpurun_out/${tag}_sweep.log
