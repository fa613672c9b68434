#!/usr/bin/env python
"""PCIe ceiling of this box: pinned H2D, D2H and both at once (GB/s), the denominators of the e2e numbers."""
import json
import time

import torch

n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


run(True, True, 1)
out = {"bytes": n, "h2d_gbs": n / run(True, False) / 1e9, "d2h_gbs": n / run(False, True) / 1e9}
t = run(True, True)
out["both_each_gbs"] = n / t / 1e9
out["both_sum_gbs"] = 2 * n / t / 1e9
print(json.dumps(out))
