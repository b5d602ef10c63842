set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "### $*" >> gpurun_out/sweep1.log; env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('clocks'))
" >> gpurun_out/sweep1.log; }
run BP_ASYNC_E=4
run BP_ASYNC_E=3
run BP_ASYNC_E=2
echo "### envs 947200 (5.0 waves of 1480)" >> gpurun_out/sweep1.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --envs 947200 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> gpurun_out/sweep1.log
echo "### envs 1136640 (6.0 waves)" >> gpurun_out/sweep1.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --envs 1136640 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> gpurun_out/sweep1.log
cat gpurun_out/sweep1.log
