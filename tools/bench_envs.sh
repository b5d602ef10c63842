cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/r1_bench_envs.jsonl
for e in GripperTouch-v0 BlocksTouch-v0 BlocksTouchCurriculum-v0 BlocksTouchChoose-v0 BlocksTouchChooseCurriculum-v0 BlocksTouchVariation-v0 ToppleTower-v0; do
  timeout 600 python bench.py --env $e --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | grep '^{' >> gpurun_out/r1_bench_envs.jsonl
done
python -c "
import json
for l in open('gpurun_out/r1_bench_envs.jsonl'):
    d=json.loads(l); e=d['episode_stats']
    print(d['config']['env_id'], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'B/step %.1f'%d['roofline']['algorithmic_bytes_per_env_step'], 'frac %.3f'%d['roofline']['frac'], 'full-physics %.3f'%(e['worker_steps']/e['steps']), 'success %.3f'%e['success_rate'])"
