#!/usr/bin/env python3
"""Sweeps the occupancy cap of the constraint interpreter (SS_CE_SMEM_PAD) on the real starknet AIR."""
import json, os, random, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import sandstorm_b200 as ss
    from sandstorm_b200.air import compile_program
    from sandstorm_b200.air.evaluate import evaluate
    from sandstorm_b200.air.layouts import load_layout
    from sandstorm_b200.air.expr import P
    from sandstorm_b200.matrix import inv_x_minus_c
    import numpy as np
    log_n = int(os.environ.get("BI_LOG_N", "20"))
    L = load_layout("starknet"); rnd = random.Random(1)
    prog = compile_program(L.composition(1 << log_n, inv_x_minus_one_col=L.num_columns), log_n, 1, [rnd.randrange(P) for _ in range(6)],
                           [rnd.randrange(P) for _ in range(17)], [rnd.randrange(P)])
    N = 2 << log_n
    g = torch.Generator(device="cuda").manual_seed(1)
    t = torch.randint(0, 2**62, (L.num_columns + 1, N, 4), dtype=torch.int64, device="cuda", generator=g); t[:, :, 3] &= (1 << 58) - 1
    m = ss.Matrix(t)
    one = np.array([0xffffffffffffffe1, 0xffffffffffffffff, 0xffffffffffffffff, 0x07fffffffffffdf0], dtype=np.uint64)
    inv_x_minus_c(t[L.num_columns], one)
    for pad in [int(x) for x in sys.argv[2].split(",")]:
        os.environ["SS_CE_SMEM_PAD"] = str(pad)
        for _ in range(2): evaluate(prog, m, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): evaluate(prog, m, 1)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"pad": pad, "ms": e0.elapsed_time(e1) / 3, "ns_per_row": e0.elapsed_time(e1) / 3 * 1e6 / N, "slots": prog.n_slots, "mul": prog.n_mul}), flush=True)
else:
    # the program compile (tables) takes ~25 s: do it once in the child and sweep there
    subprocess.check_call([sys.executable, __file__, "child", "0"])
