#!/usr/bin/env python3
"""Times ss_constraint_eval on a synthetic AIR with the size profile of the starknet layout
(195 constraints, ~650 multiplications, 267 distinct taps over 10 columns with offsets up
to 33158, zerofier periods up to 32768, 19 single-point denominators) until the real AIR is transpiled."""
import json
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.air import Challenge, Constant, Hint, Periodic, Trace, X, compile_program, composition_constraint  # noqa: E402
from sandstorm_b200.air.evaluate import evaluate  # noqa: E402
from sandstorm_b200.air.expr import P  # noqa: E402


def synthetic_air(n, n_cols=10, n_constraints=195, seed=1):
    rnd = random.Random(seed)
    g = pow(3, (P - 1) // n, P)
    one = Constant(1)
    periods = [k for k in (1, 2, 4, 8, 16, 64, 128, 256, 512, 1024, 16384, 32768) if k <= n]
    zinv = {k: one / (X.pow(n // k) - one) for k in periods}
    points = [X - Constant(pow(g, rnd.randrange(n), P)) for _ in range(19)]
    per = [Periodic([rnd.randrange(P) for _ in range(min(512, n // 4))], min(2048, n)) for _ in range(2)]
    taps = [(rnd.randrange(n_cols), rnd.choice([0, 1, 2, 3, 4, 8, 16, 31, 64, 127, 255, 390, 2045, 16775, 32763, 33158]) % n) for _ in range(267)]
    cons = []
    for i in range(n_constraints):
        a, b, c, d = (Trace(*rnd.choice(taps)) for _ in range(4))
        body = a * b - c * Constant(rnd.randrange(P)) + d
        if i % 3 == 0:
            body = body * (Trace(*rnd.choice(taps)) - Challenge(i % 6)) + Hint(i % 17)
        if i % 7 == 0:
            body = body * per[i % 2] - Trace(*rnd.choice(taps))
        if i % 10 == 0:
            cons.append(body / points[(i // 10) % 19])
        elif i % 11 == 0:
            cons.append(body * points[i % 19] * zinv[rnd.choice(periods)])
        else:
            cons.append(body * zinv[rnd.choice(periods)])
    return cons


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    log_b = 1
    n, N = 1 << log_n, 1 << (log_n + log_b)
    rnd = random.Random(2)
    t0 = time.time()
    expr = composition_constraint(synthetic_air(n))
    prog = compile_program(expr, log_n, log_b, [rnd.randrange(P) for _ in range(6)], [rnd.randrange(P) for _ in range(17)], [rnd.randrange(P)])
    t_compile = time.time() - t0
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    lde = torch.randint(0, 2**62, (10, N, 4), dtype=torch.int64, device="cuda", generator=g)
    lde[:, :, 3] &= (1 << 58) - 1
    m = ss.Matrix(lde)
    for _ in range(2):
        evaluate(prog, m, log_b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        evaluate(prog, m, log_b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rec = {"log_n": log_n, "rows": N, "ms": ms, "ns_per_row": ms * 1e6 / N, "n_instr": prog.n_instr, "n_mul": prog.n_mul, "n_addsub": prog.n_addsub,
           "taps": prog.n_trace_taps, "tables": prog.table_sizes, "slots": prog.n_slots, "batch_inv": prog.n_batch_inv, "compile_s": round(t_compile, 2),
           "field_ops_per_s": (prog.n_mul + prog.n_addsub) * N / (ms * 1e-3), "algo_GBps": 11 * 32 * N / (ms * 1e-3) / 1e9}
    print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rec, open("gpurun_out/bench_constraints.json", "w"), indent=1)


if __name__ == "__main__":
    main()
