"""HBM calibration on this box: pure write (fill), copy, pure read (sum) bandwidth with torch."""
import json

import torch

n = 1 << 30  # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
out = {}


def t(fn, bytes_, name, k=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = round(bytes_ / best / 1e6, 1)


t(lambda: a.fill_(1.0), 4 * n, "fill_GBps")
t(lambda: b.copy_(a), 8 * n, "copy_GBps")
t(lambda: a.sum(), 4 * n, "read_sum_GBps")
print(json.dumps(out))
