#!/usr/bin/env python3
"""Stage timing of the device hot path for a layout / trace length (synthetic trace)."""
import json
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.prover import HotPathProver  # noqa: E402

layout, log_n = sys.argv[1], int(sys.argv[2])
torch.cuda.set_device(0)
hp = HotPathProver(layout, log_n)
L = hp.layout
g = torch.Generator(device="cuda").manual_seed(1)


def cols(c):
    t = torch.randint(0, 2**62, (c, 1 << log_n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, :, 3] &= (1 << 58) - 1
    return ss.Matrix(t)


base, ext = cols(L.num_base_columns), cols(L.num_extension_columns)
t0 = time.time(); hp.composition_template(); t_compile = time.time() - t0
for rep in range(2):
    hp.timeline.clear()
    t0 = time.time()
    res = hp.prove(base, ext)
    torch.cuda.synchronize()
    wall = time.time() - t0
st = hp.stage_ms()
print(json.dumps({"layout": layout, "log_n": log_n, "compile_s": round(t_compile, 1), "wall_s": round(wall, 3), "stages_ms": {k: round(v, 2) for k, v in st.items()},
                  "gpu_ms": round(sum(st.values()), 1), "fri_layers": len(res.fri_roots), "peak_GB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}))
