#!/usr/bin/env python3
"""Times ss_inv_x_minus_c (and checks a few values against big ints).  Usage: python tools/bench_inv.py [log_n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.air.expr import P  # noqa: E402
from sandstorm_b200.matrix import inv_x_minus_c  # noqa: E402

R = 2**256
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
N = 1 << log_n
out = torch.empty((N, 4), dtype=torch.int64, device="cuda")
c = 12345678901234567890123456789
cm = np.array([((c * R % P) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
for _ in range(2):
    inv_x_minus_c(out, cm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    inv_x_minus_c(out, cm)
e1.record()
torch.cuda.synchronize()
w = pow(3, (P - 1) // N, P)
got = out.cpu().numpy().view(np.uint64)
for i in (0, 1, 4095, 4096, N // 2 + 7, N - 1):
    v = (int(got[i][0]) | int(got[i][1]) << 64 | int(got[i][2]) << 128 | int(got[i][3]) << 192) * pow(R, -1, P) % P
    assert v == pow(3 * pow(w, i, P) - c, -1, P), i
print({"log_n": log_n, "ms": e0.elapsed_time(e1) / 5})
