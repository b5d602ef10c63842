cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep6.log
run() { env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']; n=d['steps']*8192
print('$*', '%.4g'%d['value'], 'iters/slab %.1f passes/slab %.1f fill %.1f'%(e['sched_iterations']/n, e['sched_passes']/n, e['worker_steps']/max(e['sched_passes'],1)))" >> gpurun_out/sweep6.log; }
run BP_NO_STEAL=1
run BP_NO_STEAL=0
run BP_NO_STEAL=1
run BP_NO_STEAL=0
cat gpurun_out/sweep6.log
