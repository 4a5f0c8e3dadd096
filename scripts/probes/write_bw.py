"""Probe: write-only HBM bandwidth on this GPU (the slot-grid kernel only writes), beside the read+write copy figure of
MEASURED_PEAKS.json. torch ops only: fill_ (elementwise store kernel), zero_ (memset), copy_ (read + write)."""
import json
import torch

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

out = {}
for gib in (1, 4):
    n = gib * (1 << 30) // 4
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    y = torch.empty(n, dtype=torch.int32, device="cuda")
    out["fill_%dGiB_GBs" % gib] = n * 4 / 1e6 / timeit(lambda: x.fill_(-1))
    out["memset_%dGiB_GBs" % gib] = n * 4 / 1e6 / timeit(lambda: x.zero_())
    out["copy_%dGiB_GBs_rw" % gib] = 2 * n * 4 / 1e6 / timeit(lambda: y.copy_(x))
    out["read_sum_%dGiB_GBs" % gib] = n * 4 / 1e6 / timeit(lambda: x.sum())
    del x, y
print(json.dumps(out))
