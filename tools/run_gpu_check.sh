# usage: bash tools/run_gpu_check.sh <tag>  -- GPU parity tests + short bench of the duo and async kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
for mode in duo async; do
  BP_STEP_KERNEL=$mode timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep '^{' > gpurun_out/${tag}_bench_$mode.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/${tag}_bench_$mode.json')); print('$mode', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
