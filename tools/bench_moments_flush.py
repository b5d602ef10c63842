import os, sys, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import blockpuzzle_gym_b200 as bpg
dev = torch.device("cuda", 0)
x = torch.randn(1 << 20, 40, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush2 = torch.empty(64 << 20, dtype=torch.float32, device=dev)
nz = bpg.Normalizer(40)
def run(mode):
    for _ in range(3): nz.update(x)
    ts = []
    for _ in range(20):
        if mode == "write": flush.zero_()
        elif mode == "write+read": flush.zero_(); flush2.sum()
        elif mode == "read": flush2.sum()
        torch.cuda._sleep(400000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nz.update(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(mode, "%.1f us" % (ms * 1e3), "%.1f %%" % (100 * x.numel() * 4 / (ms * 1e-3) / 1e9 / 6536.7))
for m in ("write", "write+read", "read", "none"): run(m)
