"""usage: python tools/bench_moments.py  -- Normalizer.update ([1 Mi, 40] float32) launch time for the blocks-per-SM setting in BP_MOMENTS_BLOCKS_PER_SM"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockpuzzle_gym_b200 as bpg
dev = torch.device("cuda", 0)
x = torch.randn(1 << 20, 40, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nz = bpg.Normalizer(40)
for _ in range(3): nz.update(x)
ts = []
for _ in range(20):
    flush.zero_(); torch.cuda._sleep(400000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); nz.update(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
print(os.environ.get("BP_MOMENTS_BLOCKS_PER_SM", "default"), "%.1f us" % (ms * 1e3), "%.1f %% of 6536.7 GB/s" % (100 * x.numel() * 4 / (ms * 1e-3) / 1e9 / 6536.7))
