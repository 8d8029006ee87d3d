#!/usr/bin/env python3
"""Goldilocks NTT timing on one GPU (run on the GPU box): forward DIF (natural -> bit-reversed) + inverse DIT (bit-reversed ->
natural, with its 1/n sweep), averaged, for 2^20 x 64, 2^24 x 8 and 2^26 x 4 columns.  Prints field-ops/s (1.5 n log n per
transform) and the read-once / write-once GB/s the roofline fraction in DESIGN.md §4.7 is quoted on.
    gpurun -- python tools/bench_goldilocks.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss
from sandstorm_b200 import goldilocks as glk
for log_n, cols in ((20, 64), (24, 8), (26, 4)):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, 2**62, (cols, 1 << log_n), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(2): glk.ntt_(a); glk.ntt_(a, inverse=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps): glk.ntt_(a, in_order=ss.ORDER_NATURAL, out_order=ss.ORDER_BITREV); glk.ntt_(a, inverse=True, in_order=ss.ORDER_BITREV, out_order=ss.ORDER_NATURAL)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (2 * reps)
    n = 1 << log_n
    ops = cols * 1.5 * n * log_n
    print(f"2^{log_n} x {cols}: {ms:.3f} ms per transform batch, {ops / ms / 1e6:.1f} Gfield-ops/s, {cols * n * 16 / ms / 1e6:.1f} GB/s algorithmic (read once + write once)")
