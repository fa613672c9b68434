#!/usr/bin/env python
"""Stream ceilings of this GPU's HBM, measured with library kernels (context for the roofline fractions):
read-only (sum), write-only (fill), copy 1:1.  GB/s of bytes actually moved, CUDA events, best of 10."""
import json

import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.int32, device="cuda").random_()
b = torch.empty_like(a)


def best(fn, nbytes, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / (min(t) * 1e-3) / 1e9


print(json.dumps({"read_only_sum_gbs": best(lambda: a.sum(), 4 * n), "write_only_fill_gbs": best(lambda: b.fill_(7), 4 * n),
                  "copy_gbs": best(lambda: b.copy_(a), 8 * n)}))
