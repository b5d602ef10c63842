# usage: bash tools/scale.sh <N> <tag>   -- the driver's N-GPU launch of bench.py
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; N=$1; tag=${2:-r1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_${N}gpu.json') if l.startswith('{')][-1]; print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/${tag}_bench_${N}gpu.err
