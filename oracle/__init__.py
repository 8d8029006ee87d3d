"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liboracle.so`` (the plain-C CPU restatement of the
reference's hot path).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; nothing under ``sandstorm_b200/`` does.

Elements are numpy ``uint64`` arrays of shape ``(..., 4)``: little-endian limbs
of the Montgomery form x*2^256 mod p (reference: crypto/src/utils.rs:15-17).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

P = 2**251 + 17 * 2**192 + 1
R = 2**256
R_INV = pow(R, -1, P)
GENERATOR = 3

TREE_KECCAK, TREE_KECCAK_M20, TREE_FRIENDLY, TREE_BLAKE2S_M20, TREE_SHA256 = range(5)
HASH_KECCAK, HASH_KECCAK_M20, HASH_BLAKE2S, HASH_BLAKE2S_M20, HASH_SHA256 = range(5)


def build(force: bool = False) -> str:
    """Compile liboracle.so in-tree (gcc, seconds)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE)
        if f.endswith((".c", ".h"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------- conversions
def to_mont(values) -> np.ndarray:
    """Python ints (canonical) -> uint64[...,4] Montgomery limbs."""
    vals = [int(v) % P * R % P for v in values]
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_mont(arr: np.ndarray) -> list[int]:
    """uint64[...,4] Montgomery limbs -> Python ints (canonical)."""
    flat = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in flat:
        v = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        out.append(v * R_INV % P)
    return out


def random_felts(rng: np.random.Generator, *shape: int) -> np.ndarray:
    """Uniform-ish elements in [0,p) already in Montgomery form (any residue is a
    valid Montgomery representative, so sampling the limbs directly is uniform)."""
    n = int(np.prod(shape))
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x07FFFFFFFFFFFFFF)     # < 2^251 < p
    return a.reshape(*shape, 4)


# ------------------------------------------------------------------------ field
def fp_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    f = lib().fp_mul
    for i in range(a.shape[0]):
        f(_ptr(out[i]), _ptr(a[i]), _ptr(b[i]))
    return out


# -------------------------------------------------------------------------- NTT
def ntt(cols: np.ndarray, inverse: bool = False) -> np.ndarray:
    """cols: uint64[n_cols, n, 4]; per-column Radix2EvaluationDomain fft / ifft."""
    a = np.array(cols, dtype=np.uint64, order="C", copy=True)
    n_cols, n, _ = a.shape
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().oracle_ntt_fp252_batch(_ptr(a), ctypes.c_int(n_cols), ctypes.c_int(log_n), ctypes.c_int(int(inverse)))
    return a


def coset_ntt(col: np.ndarray, inverse: bool = False) -> np.ndarray:
    """one column uint64[n, 4]: ark-poly coset_fft / coset_ifft with offset 3."""
    a = np.array(col, dtype=np.uint64, order="C", copy=True)
    lib().oracle_coset_ntt_fp252(_ptr(a), ctypes.c_int(a.shape[0].bit_length() - 1), ctypes.c_int(int(inverse)))
    return a


def constraint_eval(blob: bytes, cols: np.ndarray, log_N: int, out: np.ndarray | None = None, rows=None, log_step: int = 0) -> np.ndarray:
    """cols: uint64[n_cols, N, 4] (column-major LDE matrix).  Evaluates a program blob (sandstorm_b200/air/program.py) on the
    rows begin + (k << log_step), k < count; result k is stored at out[row >> log_step]."""
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    N = 1 << log_N
    assert a.shape[1] == N
    if out is None:
        out = np.zeros((N >> log_step, 4), dtype=np.uint64)
    begin, count = rows if rows is not None else (0, 0)
    rc = lib().oracle_constraint_eval(blob, ctypes.c_size_t(len(blob)), _ptr(a), ctypes.c_uint64(N), ctypes.c_int(log_N),
                                      ctypes.c_uint64(begin), ctypes.c_uint64(count), ctypes.c_int(log_step), _ptr(out))
    if rc:
        raise ValueError(f"oracle_constraint_eval: bad program blob ({rc})")
    return out


def inv_x_minus_c(log_N: int, c: int, log_step: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """out[i] = 1 / (3 w_N^i - c) on the rows that are multiples of 2^log_step (c: canonical int)."""
    if out is None:
        out = np.zeros((1 << log_N, 4), dtype=np.uint64)
    lib().oracle_inv_x_minus_c(ctypes.c_int(log_N), ctypes.c_int(log_step), _ptr(to_mont([c])), _ptr(out))
    return out


def horner(coeffs: np.ndarray, z: int) -> int:
    """P(z) for natural-order Montgomery coefficients uint64[n, 4]; canonical int."""
    a = np.ascontiguousarray(coeffs, dtype=np.uint64)
    r = np.zeros((1, 4), dtype=np.uint64)
    lib().oracle_horner(_ptr(a), ctypes.c_uint64(a.shape[0]), _ptr(to_mont([z])), _ptr(r))
    return from_mont(r)[0]


def fri_fold(evals: np.ndarray, log_fold: int, alpha: int, offset: int, starkware_scale: bool = True) -> np.ndarray:
    """natural-order evaluations on offset<w> -> natural-order folded evaluations; starkware_scale: no 1/F factor."""
    a = np.ascontiguousarray(evals, dtype=np.uint64)
    log_n = a.shape[0].bit_length() - 1
    out = np.zeros((a.shape[0] >> log_fold, 4), dtype=np.uint64)
    lib().oracle_fri_fold(_ptr(a), ctypes.c_int(log_n), ctypes.c_int(log_fold), _ptr(to_mont([alpha])), _ptr(to_mont([offset])),
                          ctypes.c_int(int(starkware_scale)), _ptr(out))
    return out


def lde(cols: np.ndarray, log_blowup: int) -> np.ndarray:
    """Matrix::interpolate then Matrix::evaluate on the coset 3*<w_N>."""
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    n_cols, n, _ = a.shape
    log_n = n.bit_length() - 1
    out = np.empty((n_cols, n << log_blowup, 4), dtype=np.uint64)
    lib().oracle_lde_fp252_batch(_ptr(a), ctypes.c_int(n_cols), ctypes.c_int(log_n), ctypes.c_int(log_blowup), _ptr(out))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core."""
    lib().oracle_set_threads(ctypes.c_int(int(n)))


# ----------------------------------------------------------------------- hashes
def hash_bytes(kind: int, data: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().oracle_hash_bytes(ctypes.c_int(kind), data, ctypes.c_size_t(len(data)), out)
    return out.raw


def pedersen_hash(a: int, b: int) -> int:
    am, bm = to_mont([a]), to_mont([b])
    out = np.empty((1, 4), dtype=np.uint64)
    lib().oracle_pedersen_hash(_ptr(out), _ptr(am), _ptr(bm))
    return from_mont(out)[0]


def pedersen_hash_mont(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    f = lib().oracle_pedersen_hash
    for i in range(a.shape[0]):
        f(_ptr(out[i]), _ptr(a[i]), _ptr(b[i]))
    return out


def hash_rows(kind: int, cols: np.ndarray, bitrev_rows: bool = False) -> np.ndarray:
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    n_cols, n, _ = a.shape
    out = np.empty((n, 32), dtype=np.uint8)
    lib().oracle_hash_rows(ctypes.c_int(kind), _ptr(a), ctypes.c_int(n_cols), ctypes.c_size_t(n), ctypes.c_int(int(bitrev_rows)), _ptr(out))
    return out


def merkle_build(kind: int, cols: np.ndarray, n_friendly: int = 22, bitrev_rows: bool = False):
    """Returns (nodes uint8[n,32] with slot 0 unused, leaves uint8[n,32], root_bytes)."""
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    n_cols, n, _ = a.shape
    log_rows = n.bit_length() - 1
    nodes = np.zeros((n, 32), dtype=np.uint8)
    leaves = np.zeros((n, 32), dtype=np.uint8)
    rc = lib().oracle_merkle_build(ctypes.c_int(kind), ctypes.c_int(n_friendly), _ptr(a), ctypes.c_int(n_cols),
                                   ctypes.c_int(log_rows), ctypes.c_int(int(bitrev_rows)), _ptr(nodes), _ptr(leaves))
    if rc != 0:
        raise ValueError(f"oracle_merkle_build failed: {rc}")
    root = ctypes.create_string_buffer(32)
    lib().oracle_merkle_root_bytes(ctypes.c_int(kind), ctypes.c_int(n_friendly), ctypes.c_int(n_cols), ctypes.c_int(log_rows), _ptr(nodes), root)
    return nodes, leaves, root.raw


def merkle_find_index(hash_kind: int, leaf: bytes, siblings: list, root: bytes) -> int:
    """index of the opening (leaf, siblings leaf-level-first) under `root`, or -1 (see oracle_merkle_find_index)."""
    f = lib().oracle_merkle_find_index
    f.restype = ctypes.c_longlong
    return int(f(ctypes.c_int(hash_kind), leaf, b"".join(siblings), ctypes.c_int(len(siblings)), root))


# ----------------------------------------------------------------- Goldilocks (p = 2^64 - 2^32 + 1), stored words x * 2^64 mod p
GL_P = 2**64 - 2**32 + 1
GL_GENERATOR = 7


def gl_ntt(cols: np.ndarray, inverse: bool = False, coset: bool = False) -> np.ndarray:
    """cols: uint64[n_cols, n] stored words.  Natural order in and out (ark-poly fft / ifft / coset forms)."""
    a = np.ascontiguousarray(cols, dtype=np.uint64).copy()
    n_cols, n = a.shape
    lib().oracle_gl_ntt_cols(_ptr(a), ctypes.c_size_t(n_cols), ctypes.c_int(n.bit_length() - 1), ctypes.c_int(int(inverse)), ctypes.c_int(int(coset)))
    return a


def gl_lde(cols: np.ndarray, log_blowup: int) -> np.ndarray:
    a = np.ascontiguousarray(cols, dtype=np.uint64)
    n_cols, n = a.shape
    out = np.empty((n_cols, n << log_blowup), dtype=np.uint64)
    lib().oracle_gl_lde_cols(_ptr(a), ctypes.c_size_t(n_cols), ctypes.c_int(n.bit_length() - 1), ctypes.c_int(log_blowup), _ptr(out))
    return out
