"""Read-only / write-only / copy bandwidth with torch ops (context for the write-heavy kernels' roofline fractions)."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 29          # 512 Mi float32 = 2 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record(); fn(); y.record(); torch.cuda.synchronize()
        best = min(best, x.elapsed_time(y))
    return best * 1e-3


res = {"write_fill_GBs": 4 * n / t(lambda: a.fill_(1.0)) / 1e9,
       "read_sum_GBs": 4 * n / t(lambda: a.sum()) / 1e9,
       "copy_GBs": 8 * n / t(lambda: b.copy_(a)) / 1e9}
print(json.dumps(res))
