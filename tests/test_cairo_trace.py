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


# ---- starknet layout: the reference's example/bootloader (131072 steps, n = 2^21; every builtin of the layout is exercised: two real
# Pedersen hashes and 36 real range checks from the private input, the rest the reference's dummy instances) ------------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def bootloader():
    tr = cairo.load_bootloader(os.path.join(GOLDEN, "bootloader"), os.path.join(GOLDEN, "poseidon_params.json"))
    rnd = random.Random(0xCA2)
    challenges = [rnd.randrange(P) for _ in range(6)]
    return tr, challenges, tr.build_extension_columns(challenges), tr.gen_hints(challenges)


def test_starknet_builder_invariants(bootloader):
    tr, challenges, ext, hints = bootloader
    pi, n = tr.public_input, tr.trace_len
    assert n == 1 << 21 and len(tr.base_columns) == 9 and len(ext) == 1 and len(hints) == 17
    assert (tr.range_check_min, tr.range_check_max) == (pi.rc_min, pi.rc_max)
    assert ext[0][n - 2] == hints[4] and ext[0][n - 3] == 1 and ext[0][n - 1] == 1 and ext[0][n - 8 + 3] == hints[10]
    # the ECDSA dummy is a valid signature under the reference's own rules, the EC-op dummy is P0 + G
    sig = cairo.ecdsa_dummy_trace()
    assert sig["pubkey"] == cairo.EC_GEN and sig["message"] == cairo.pedersen_hash(1, 0)
    assert cairo.ec_op_dummy_trace()["r"] == cairo.ec_add(cairo.PEDERSEN_POINTS[0], cairo.EC_GEN)


def test_every_starknet_constraint_vanishes_on_its_zerofier(bootloader):
    """All 195 constraints of layouts/src/starknet/air.rs, as transpiled into sandstorm_b200/air/layouts/starknet.json, on a real
    trace.  Rows: one whole ECDSA period (32768 rows, the longest) at each end of the trace plus random ones; each distinct zerofier
    is solved once over those rows, then every constraint is evaluated (as a fraction, nothing divided) where its zerofier vanishes."""
    tr, challenges, ext, hints = bootloader
    L = load_layout("starknet")
    n, log_n = tr.trace_len, 21
    cols = tr.base_columns + ext
    constraints = L.constraints(n)
    assert len(constraints) == 195
    g = pow(3, (P - 1) // n, P)
    rnd = random.Random(7)
    span = 32768 + 600
    xs, x = {}, 1
    for i in range(span):
        xs[i] = x
        x = x * g % P
    x = pow(g, n - span, P)
    for i in range(n - span, n):
        xs[i] = x
        x = x * g % P
    for _ in range(300):
        i = rnd.randrange(n)
        xs[i] = pow(g, i, P)
    rows = sorted(xs)
    zero_rows = {}

    def vanishing(zf):
        if zf not in zero_rows:
            zero_rows[zf] = [i for i in rows if eval_fraction(zf, i, cols, log_n, challenges, hints, {}, x=xs[i])[0] == 0]
        return zero_rows[zf]

    for k, c in enumerate(constraints):
        cand = sorted(set().union(*[vanishing(z) for z in divisors(c)]))
        if len(cand) > 900:
            cand = cand[:300] + cand[-300:] + random.Random(k).sample(cand, 300)
        hits = 0
        for i in cand:
            num, den = eval_fraction(c, i, cols, log_n, challenges, hints, {}, x=xs[i])
            if den != 0:
                continue
            hits += 1
            assert num == 0, f"starknet constraint {k} is violated at trace row {i}"
        assert hits > 0, f"starknet constraint {k}: no sampled row lies on its zerofier"


def test_a_wrong_starknet_trace_is_caught(bootloader):
    """one flipped cell in each builtin's region (Poseidon full-round state, ECDSA doubling slope, EC-op partial sum, bitwise diluted
    chunk, Pedersen suffix) must violate some constraint on the instance's rows."""
    tr, challenges, ext, hints = bootloader
    L = load_layout("starknet")
    constraints = L.constraints(tr.trace_len)
    for col, row in ((8, 64 + cairo.SN_POSEIDON["Full1"]), (8, 128 + cairo.SN_ECDSA["PubkeyDoublingSlope"]), (8, 64 + cairo.SN_ECOP["RPartialSumX"]),
                     (7, 1 + 16), (3, 5)):
        cols = list(tr.base_columns) + ext
        cols[col] = list(cols[col])
        cols[col][row] = (cols[col][row] + 1) % P
        near = sorted({(row // p) * p + d for p in (1, 8, 16, 64, 128, 256, 512) for d in (0, -p) if (row // p) * p + d >= 0})
        bad = 0
        for c in constraints:
            for i in near:
                num, den = eval_fraction(c, i, cols, 21, challenges, hints, {})
                bad += den == 0 and num != 0
        assert bad > 0, (col, row)
