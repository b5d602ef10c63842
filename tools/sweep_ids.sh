# usage: bash tools/sweep_ids.sh <tag> "VAR=val ..." ...  -- bench.py --env <id> for the non-headline ids under each environment setting
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; shift
: > gpurun_out/${tag}_ids.log
for id in GripperTouch-v0 BlocksTouchChoose-v0 BlocksTouchVariation-v0 ToppleTower-v0; do
for setting in "$@"; do
  env $setting timeout 300 python bench.py --env $id --steps 4 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$id', '$setting', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'])" >> gpurun_out/${tag}_ids.log
done; done
cat gpurun_out/${tag}_ids.log
