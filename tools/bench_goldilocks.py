#!/usr/bin/env python3
"""Goldilocks NTT microbench (BASELINE config 5 shape for the single-limb field): batched forward coset NTT,
field-ops/s and achieved HBM GB/s (algorithmic bytes = 2 * 8 B per element per pass).  Usage: bench_goldilocks.py [log_n ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200 import goldilocks as glk  # noqa: E402

for log_n in [int(v) for v in sys.argv[1:]] or [16, 20, 24, 26]:
    n_cols = max(1, min(64, (1 << 29) >> (log_n + 3)))          # ~512 MiB per batch
    g = torch.Generator(device="cuda").manual_seed(log_n)
    a = torch.randint(0, 2**62, (n_cols, 1 << log_n), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(2):
        glk.ntt_(a, coset=False, out_order=ss.ORDER_BITREV)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        glk.ntt_(a, coset=False, out_order=ss.ORDER_BITREV)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    passes = 1 if log_n <= 12 else 1 + -(-(log_n - 12) // 9)
    ops = 1.5 * (1 << log_n) * log_n * n_cols
    rec = {"field": "goldilocks", "log_n": log_n, "n_cols": n_cols, "ms": round(ms, 4), "field_ops_per_s": ops / (ms * 1e-3), "passes": passes,
           "algo_GBps": n_cols * (1 << log_n) * 16 * passes / (ms * 1e-3) / 1e9}
    print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_goldilocks.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
