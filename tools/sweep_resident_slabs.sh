cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/sweep_resident_slabs.log
run() { env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', '%.4g'%d['value'])" >> gpurun_out/sweep_resident_slabs.log; }
run BP_SMEM_PAD=0
run BP_SMEM_PAD=2900
run BP_SMEM_PAD=6100
run BP_SMEM_PAD=10200
cat gpurun_out/sweep_resident_slabs.log
