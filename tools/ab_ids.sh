# usage: bash tools/ab_ids.sh <tag> "id1 id2 ..." libA.so libB.so ...  -- bench.py --env <id> on several builds of the library, same box
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=$1; ids=$2; shift; shift
: > gpurun_out/${tag}_abids.log
for id in $ids; do for lib in "$@"; do
  BP_LIB_PATH=$PWD/$lib timeout 300 python bench.py --env $id --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$id', '$lib', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'])" >> gpurun_out/${tag}_abids.log
done; done
cat gpurun_out/${tag}_abids.log
