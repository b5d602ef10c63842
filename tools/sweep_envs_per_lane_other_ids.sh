cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep_envs_per_lane_other_ids.log
for e in BlocksTouchChoose-v0 ToppleTower-v0 BlocksTouchVariation-v0 GripperTouch-v0; do for E in 4 3 2; do
  BP_ASYNC_E=$E timeout 600 python bench.py --env $e --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e E=$E', '%.4g'%d['value'])" >> gpurun_out/sweep_envs_per_lane_other_ids.log
done; done; cat gpurun_out/sweep_envs_per_lane_other_ids.log
