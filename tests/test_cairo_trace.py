"""The reference's own example (example/trace.bin, memory.bin, air-public-input.json: array-sum, recursive layout,
16384 steps) through the restated trace builder (oracle/cairo.py) and the TRANSPILED AIR: every one of the 93 constraints
of layouts/src/recursive/air.rs must vanish on its zerofier set.  Random columns can never show that; this pins the
meaning of sandstorm_b200/air/layouts/recursive.json (signs, column indices, strides, periodic columns, hints), not
just its shape."""
import os
import random

import pytest

from oracle import cairo
from sandstorm_b200.air.expr import P
from sandstorm_b200.air.layouts import load_layout
from air_ref import divisors, eval_fraction

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "array_sum")


@pytest.fixture(scope="module")
def example():
    tr = cairo.load_example(FIXTURE)
    rnd = random.Random(0xCA1)
    challenges = [rnd.randrange(P) for _ in range(6)]
    return tr, challenges, tr.build_extension_columns(challenges), tr.gen_hints(challenges)


def test_fixture_shape_and_builder_invariants(example):
    tr, challenges, ext, hints = example
    pi = tr.public_input
    assert tr.trace_len == 1 << 18 and len(tr.base_columns) == 7 and len(ext) == 3
    assert (tr.range_check_min, tr.range_check_max) == (pi.rc_min, pi.rc_max)              # air-public-input.json
    assert tr.initial_registers == (45, 45, 1) and tr.final_registers == (76, 45, 5)
    n = tr.trace_len
    # the permutation products close: memory ends at the public-memory quotient, range check and diluted check at 1
    assert ext[2][n - 2] == hints[4] and ext[2][n - 3] == 1 and ext[1][n - 1] == 1
    assert ext[0][n - 1] == hints[10]                                                       # diluted cumulative value


def test_every_constraint_vanishes_on_its_zerofier(example):
    tr, challenges, ext, hints = example
    L = load_layout("recursive")
    n, log_n = tr.trace_len, 18
    cols = tr.base_columns + ext
    constraints = L.constraints(n)
    assert len(constraints) == 93
    rnd = random.Random(7)
    rows = list(range(0, 2100)) + list(range(n - 2100, n)) + [rnd.randrange(n) for _ in range(800)]
    for k, c in enumerate(constraints):
        hits, zs = 0, divisors(c)
        for i in rows:
            # (cheap pre-filter: the zerofiers are tiny x-only trees; the whole constraint is evaluated only where one vanishes)
            if all(eval_fraction(zf, i, cols, log_n, challenges, hints, {})[0] != 0 for zf in zs):
                continue
            num, den = eval_fraction(c, i, cols, log_n, challenges, hints, {})
            assert den == 0
            hits += 1
            assert num == 0, f"constraint {k} is violated at trace row {i}"
        assert hits > 0, f"constraint {k}: no sampled row lies on its zerofier"


def test_a_wrong_trace_is_caught(example):
    """sanity of the check itself: flip one cell, some constraint must fail."""
    tr, challenges, ext, hints = example
    L = load_layout("recursive")
    cols = [list(c) for c in tr.base_columns] + ext
    cols[6][16 + cairo.AUXILIARY["Res"]] = (cols[6][16 + cairo.AUXILIARY["Res"]] + 1) % P
    bad = 0
    for c in L.constraints(tr.trace_len):
        for i in (0, 16, 32):
            num, den = eval_fraction(c, i, cols, 18, challenges, hints, {})
            bad += den == 0 and num != 0
    assert bad > 0
