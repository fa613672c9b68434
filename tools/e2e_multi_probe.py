#!/usr/bin/env python
"""End to end through the single-process N-device layer (bn_multi_encode / bn_multi_decode): one process, pinned host
buffers holding n_devices x 1e9 bases, every H2D / D2H copy inside the timed region.  One JSON line per device count."""
from __future__ import annotations

import json
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import bitnuc_b200 as bn
from bitnuc_b200 import device as dv
from bitnuc_b200.multi import MultiContext

PER = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000_000
counts = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [torch.cuda.device_count()]
for n_dev in counts:
    for chunk_mib in (64, 256):
        m = MultiContext(n_dev, reduce="p2p")
        m.set_chunk_bytes(chunk_mib << 20)
        n = PER * n_dev
        ctx0 = m.contexts[0]
        h_seq = ctx0.pinned_empty(n, np.uint8)
        for i in range(n_dev):  # every shard generated on its own device
            with torch.cuda.device(m.devices[i]):
                h_seq[i * PER:(i + 1) * PER] = dv.synth_ascii(0x5EEDB17C0DE5, 0, i * PER, PER, device=torch.device("cuda", m.devices[i])).cpu().numpy()
        h_words = ctx0.pinned_empty(dv.words_for(n), np.uint64)
        h_back = ctx0.pinned_empty(n, np.uint8)
        m.encode_np(h_seq, out=h_words)
        m.decode_np(h_words, n, out=h_back)
        steps = 4
        t0 = time.perf_counter()
        for _ in range(steps):
            m.encode_np(h_seq, out=h_words)
            m.decode_np(h_words, n, out=h_back)
        dt = (time.perf_counter() - t0) / steps
        ok = bool(np.array_equal(h_back, h_seq))
        print(json.dumps({"probe": "bn_multi_e2e_serial", "n_gpus": n_dev, "chunk_mib": chunk_mib, "bases": n, "ms_per_step": dt * 1e3,
                          "gbases_s": 2 * n / dt / 1e9, "pcie_gbs_aggregate": 2 * 1.25 * n / dt / 1e9, "ok": ok}), flush=True)
        del h_seq, h_words, h_back
        m.close()
