#!/usr/bin/env python3
"""Times ss_constraint_eval on the REAL starknet AIR composition and on the starknet DEEP quotient
(269 taps, 2 composition columns) at n = 2^log_n; prints one JSON line per program.
Usage: python tools/bench_ce.py [log_n] ; SS_CE_MINB=5|6|7 selects the occupancy variant of the kernel."""
import json
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.air import compile_program  # noqa: E402
from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms  # noqa: E402
from sandstorm_b200.air.evaluate import evaluate  # noqa: E402
from sandstorm_b200.air.expr import P  # noqa: E402
from sandstorm_b200.air.layouts import load_layout  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    log_b = 1
    n, N = 1 << log_n, 1 << (log_n + log_b)
    rnd = random.Random(2)
    L = load_layout("starknet")
    C = L.num_columns
    t0 = time.time()
    comp = compile_program(L.composition(n, inv_x_minus_one_col=C + 2), log_n, log_b, [rnd.randrange(P) for _ in range(L.n_challenges())],
                           [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)])
    g = pow(3, (P - 1) // n, P)
    tt, ct = deep_terms(L.taps(), [rnd.randrange(P) for _ in L.taps()], [rnd.randrange(P) for _ in range(2)], C, rnd.randrange(P), P)
    deep = compile_program(deep_expr_shifted(tt, ct, C + 3, C + 4, g, P), log_n, log_b)
    t_compile = time.time() - t0
    torch.cuda.set_device(0)
    gen = torch.Generator(device="cuda").manual_seed(5)
    lde = torch.randint(0, 2**62, (C + 5, N, 4), dtype=torch.int64, device="cuda", generator=gen)
    lde[:, :, 3] &= (1 << 58) - 1
    m = ss.Matrix(lde)
    out = torch.empty((N, 4), dtype=torch.int64, device="cuda")
    c = m.ctx
    variants = [int(v) for v in os.environ.get("SS_GEN_MINB_VARIANTS", "0").split(",")]
    for name, prog, aot, mb in [(nm, pg, a, v) for nm, pg in (("composition", comp), ("deep", deep)) for a, v in [(0, 0)] + [(1, v) for v in variants]]:
        c.check(c.lib.ss_set_option(c.handle, b"ce_aot", aot))
        c.check(c.lib.ss_set_option(c.handle, b"ce_aot_minb", mb))
        for _ in range(2):
            evaluate(prog, m, log_b, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            evaluate(prog, m, log_b, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec = {"program": name, "aot": int(c.lib.ss_get_option(c.handle, b"ce_last_aot", -1)), "aot_minb": mb, "log_n": log_n, "rows": N, "ms": round(ms, 3), "ns_per_row": round(ms * 1e6 / N, 3),
               "words": prog.n_instr, "n_mul": prog.n_mul, "n_addsub": prog.n_addsub, "n_red": prog.n_red, "n_dot": prog.n_dot, "taps": prog.n_trace_taps,
               "slots": prog.n_slots, "compile_s": round(t_compile, 1), "mul_per_s": prog.n_mul * N / (ms * 1e-3)}
        print(json.dumps(rec), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/bench_ce.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
