# usage: bash tools/sweep_ids_thresholds2.sh <tag>  -- second, finer pass of tools/sweep_ids_thresholds.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-thr2}; log=gpurun_out/${tag}_sweep.log; : > $log
run() { id=$1; shift; env "$@" timeout 300 python bench.py --env $id --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['episode_stats']
print('$id', '$*', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'fill %.1f'%(e['worker_steps']/max(e['sched_passes'],1)), 'iters/pass %.1f'%(e['sched_iterations']/max(e['sched_passes'],1)))" >> $log; }
for q in 6 4 3 2 1; do run GripperTouch-v0 BP_PASS_MIN=$q; done
for r in 16 8; do run GripperTouch-v0 BP_PASS_MIN=3 BP_RESET_MIN=$r; done
for q in 16 14 12 10 8; do run BlocksTouchChoose-v0 BP_PASS_MIN=$q; done
for q in 28 16 12; do run BlocksTouchChooseCurriculum-v0 BP_PASS_MIN=$q; done
for q in 20 24; do run ToppleTower-v0 BP_PASS_MIN=$q; done
cat $log
