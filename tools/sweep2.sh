cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep2.log
for t in 0 8 16 24 32; do
  BP_DUO_MIN_READY=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']; n=d['steps']*8192
print('min_ready $t', '%.4g'%d['value'], 'iters/slab %.1f passes/slab %.1f fill %.1f'%(e['sched_iterations']/n, e['sched_passes']/n, e['worker_steps']/e['sched_passes']))" >> gpurun_out/sweep2.log
done
cat gpurun_out/sweep2.log
