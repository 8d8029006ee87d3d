#!/usr/bin/env python3
"""Generates tests/golden/*.json from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, so the outputs are committed).

  periodic_coeffs.json  — the periodic-column coefficient tables whose forward NTT the
      reference's own tests compare against EC-doubling tables / round keys:
        builtins/src/pedersen/periodic.rs:71,652  (test :1184-1209)
        builtins/src/ecdsa/periodic.rs:39,332     (test :599-625)
        builtins/src/poseidon/periodic.rs:34-193  (test :242-290)
      They are also inputs of the constraint-evaluation path.
  poseidon_round_keys.json — FULL_ROUND_KEYS_{1ST,2ND}_HALF (builtins/src/poseidon/params.rs)
      expected values of the 8-point NTT KAT.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def const_arrays(path):
    src = open(os.path.join(REF, path)).read()
    out = {}
    for m in re.finditer(r"pub const (\w+): \[Fp; (\d+)\] = \[(.*?)\];", src, re.S):
        name, n, body = m.group(1), int(m.group(2)), m.group(3)
        vals = [int(v) for v in re.findall(r'Fp!\("(\d+)"\)', body)]
        assert len(vals) == n, (name, n, len(vals))
        out[name] = [hex(v) for v in vals]
    return out


def main():
    tables = {}
    for path in ("builtins/src/pedersen/periodic.rs", "builtins/src/ecdsa/periodic.rs", "builtins/src/poseidon/periodic.rs"):
        tables.update(const_arrays(path))
    with open(os.path.join(HERE, "periodic_coeffs.json"), "w") as f:
        json.dump(tables, f, separators=(",", ":"))
    print({k: len(v) for k, v in tables.items()})

    src = open(os.path.join(REF, "builtins/src/poseidon/params.rs")).read()
    keys = {}
    for m in re.finditer(r"pub const (FULL_ROUND_KEYS_\w+): \[\[Fp; 3\]; NUM_FULL_ROUNDS / 2\] = \[(.*?)\n\];", src, re.S):
        vals = [int(v) for v in re.findall(r'Fp!\(\s*"(\d+)"\s*\)', m.group(2))]
        assert len(vals) == 12, (m.group(1), len(vals))
        keys[m.group(1)] = [[hex(v) for v in vals[3 * i:3 * i + 3]] for i in range(4)]
    with open(os.path.join(HERE, "poseidon_round_keys.json"), "w") as f:
        json.dump(keys, f, separators=(",", ":"))
    print({k: len(v) for k, v in keys.items()})


if __name__ == "__main__":
    main()


def poseidon_params():
    """poseidon_params.json — round keys of the Poseidon builtin (builtins/src/poseidon/params.rs:21-483, :520-604), inputs of the
    restated starknet trace builder (oracle/cairo.py); the MDS matrix [[3,1,1],[1,-1,1],[1,1,-2]] (:15-19) is written out there."""
    src = open(os.path.join(REF, "builtins/src/poseidon/params.rs")).read()
    out = {}
    m = re.search(r"pub const PARTIAL_ROUND_KEYS_OPTIMIZED: \[Fp; NUM_PARTIAL_ROUNDS\] = \[(.*?)\n\];", src, re.S)
    out["PARTIAL_ROUND_KEYS_OPTIMIZED"] = [hex(int(v)) for v in re.findall(r'Fp!\(\s*"(-?\d+)"\s*\)', m.group(1))]
    for name in ("FULL_ROUND_KEYS_1ST_HALF", "FULL_ROUND_KEYS_2ND_HALF"):
        m = re.search(r"pub const " + name + r": \[\[Fp; 3\]; NUM_FULL_ROUNDS / 2\] = \[(.*?)\n\];", src, re.S)
        v = [int(x) for x in re.findall(r'Fp!\(\s*"(-?\d+)"\s*\)', m.group(1))]
        out[name] = [[hex(x) for x in v[3 * i:3 * i + 3]] for i in range(4)]
    m = re.search(r"pub const PARTIAL_ROUND_KEYS: \[\[Fp; 3\]; NUM_PARTIAL_ROUNDS\] = \[(.*?)\n\];", src, re.S)
    v = [int(x) for x in re.findall(r'Fp!\(\s*"(-?\d+)"\s*\)', m.group(1))]
    out["PARTIAL_ROUND_KEYS"] = [[hex(x) for x in v[3 * i:3 * i + 3]] for i in range(83)]
    with open(os.path.join(HERE, "poseidon_params.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    poseidon_params()
