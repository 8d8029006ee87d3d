#!/usr/bin/env python3
"""Experiment: does the column stride (alignment of the columns relative to each other) matter for the tap-heavy kernels?
Runs the starknet composition + DEEP programs with col_stride = N + pad rows.  Usage: bench_ce_stride.py log_n pad [pad ...]"""
import ctypes
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.air import compile_program  # noqa: E402
from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms  # noqa: E402
from sandstorm_b200.air.expr import P  # noqa: E402
from sandstorm_b200.air.layouts import load_layout  # noqa: E402

log_n = int(sys.argv[1])
pads = [int(v) for v in sys.argv[2:]] or [0]
log_b = 1
n, N = 1 << log_n, 2 << log_n
rnd = random.Random(2)
L = load_layout("starknet")
C = L.num_columns
comp = compile_program(L.composition(n, inv_x_minus_one_col=C + 2), log_n, log_b, [rnd.randrange(P) for _ in range(L.n_challenges())],
                       [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)])
g = pow(3, (P - 1) // n, P)
tt, ct = deep_terms(L.taps(), [rnd.randrange(P) for _ in L.taps()], [rnd.randrange(P) for _ in range(2)], C, rnd.randrange(P), P)
deep = compile_program(deep_expr_shifted(tt, ct, C + 3, C + 4, g, P), log_n, log_b)
c = ss.default_context()
out = torch.empty((N, 4), dtype=torch.int64, device="cuda")
for pad in pads:
    stride = N + pad
    gen = torch.Generator(device="cuda").manual_seed(5)
    buf = torch.randint(0, 2**62, (C + 5, stride, 4), dtype=torch.int64, device="cuda", generator=gen)
    buf[:, :, 3] &= (1 << 58) - 1
    for name, prog, step in (("composition", comp, 0), ("deep_subcoset", deep, 1)):
        def run():
            c.check(c.lib.ss_constraint_eval(c.handle, prog.blob, len(prog.blob), ctypes.c_void_p(buf.data_ptr()), stride, C + 5, log_n, log_b, 0, 0, step,
                                             ctypes.c_void_p(out.data_ptr()), None))
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        rows = N >> step
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"program": name, "log_n": log_n, "pad_rows": pad, "ms": round(ms, 2), "ns_per_row": round(ms * 1e6 / rows, 3)}), flush=True)
    del buf
