cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep_thresholds.log
run() { env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']; n=d['steps']*8192
print('$*', '%.4g'%d['value'], 'iters/slab %.1f passes/slab %.1f fill %.1f'%(e['sched_iterations']/n, e['sched_passes']/n, e['worker_steps']/max(e['sched_passes'],1)))" >> gpurun_out/sweep_thresholds.log; }
for r in 1 4 8 16 32; do run BP_RESET_MIN=$r; done
for q in 16 24 28; do run BP_RESET_MIN=8 BP_PASS_MIN=$q; done
cat gpurun_out/sweep_thresholds.log
