cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep_envs_per_lane.log
run() { env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', '%.4g'%d['value'])" >> gpurun_out/sweep_envs_per_lane.log; }
# E = 5, 6 need the extra instantiations in launch_step (removed again after the sweep: E = 4 stays best)
for E in 2 3 4; do run BP_ASYNC_E=$E; done
cat gpurun_out/sweep_envs_per_lane.log
