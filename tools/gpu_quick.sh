# usage: bash tools/gpu_quick.sh <tag> [pytest -k expression]  -- GPU parity tests + one short bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; tag=${1:-x}; sel=${2:-}
if [ -n "$sel" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -25 > gpurun_out/${tag}_tests.log
else timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_tests.log; fi
cat gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-her --e2e-steps 1 --e2e-fused 8 2>gpurun_out/${tag}_bench.err | grep '^{' > gpurun_out/${tag}_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
es=d['episode_stats']
print('${tag}', 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'iters/slab %.1f passes/slab %.1f fill %.1f'%(es['sched_iterations']/8192/13, es['sched_passes']/8192/13, es['worker_steps']/max(es['sched_passes'],1)))
print({k: (v.get('value'), v.get('full_physics_frac')) for k, v in d.get('workloads', {}).items()})
PY
