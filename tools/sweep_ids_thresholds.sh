# usage: bash tools/sweep_ids_thresholds.sh <tag>  -- full-physics-pass thresholds per env id (BP_PASS_MIN x BP_FILL_RULE), same box
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-thr}; log=gpurun_out/${tag}_sweep.log; : > $log
run() { id=$1; shift; env "$@" timeout 300 python bench.py --env $id --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']
print('$id', '$*', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'fill %.1f'%(e['worker_steps']/max(e['sched_passes'],1)), 'iters/pass %.1f'%(e['sched_iterations']/max(e['sched_passes'],1)))" >> $log; }
for id in GripperTouch-v0 ToppleTower-v0; do for f in 8 0; do for q in 28 16 12 8 4; do run $id BP_PASS_MIN=$q BP_FILL_RULE=$f; done; done; done
for q in 28 24 20 16; do run BlocksTouchChoose-v0 BP_PASS_MIN=$q; done
for q in 28 16 12; do run BlocksTouchVariation-v0 BP_PASS_MIN=$q; done
cat $log
