"""Shared helpers of the test-suite: instances, right-hand sides, metrics."""
from __future__ import annotations

import os

import numpy as np

from oracle import np_oracle as npo
from oracle import oracle as orc
from two_pass_lanczos_b200 import datagen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeded_b(n: int, seed: int = 42) -> np.ndarray:
    """b ~ U[0,1), the role of StdRng::seed_from_u64(42) in the reference tests (mod.rs:439-440)."""
    return np.random.default_rng(seed).random(n)


def reference_b(n: int, seed: int = 42) -> np.ndarray:
    """b of the reference's tests restated: `StdRng::seed_from_u64(42)` uniforms (two_pass_lanczos_b200/stdrng.py; ChaCha12
    core pinned on published keystreams, seed expansion / float conversion unpinned)."""
    from two_pass_lanczos_b200 import stdrng

    return stdrng.std_rng_uniform(seed, n)


B_GENERATORS = {"numpy": seeded_b, "stdrng": reference_b}


def rhs_from_const(apply, n: int) -> np.ndarray:
    """b = A * (1/sqrt(n)) 1   (src/bin/tradeoff.rs:234-236)"""
    return apply(np.full(n, 1.0 / np.sqrt(n)))


def oracle_op(inst: datagen.KKTInstance) -> orc.SparseColMat:
    m, p = inst.m, inst.p
    j = np.arange(m, dtype=np.uint64)
    t = inst.tail.astype(np.uint64)
    h = inst.head.astype(np.uint64)
    rows = np.concatenate([j, m + t, m + h, j, j])
    cols = np.concatenate([j, j, j, m + t, m + h])
    ones = np.ones(m)
    vals = np.concatenate([inst.d, ones, -ones, ones, -ones])
    return orc.SparseColMat.try_new_from_triplets(m + p, m + p, rows, cols, vals)


def project_out_null(x: np.ndarray, m: int, p: int) -> np.ndarray:
    """(I - z z^T) x with z = [0; 1_p]/sqrt(p): every KKT matrix here is singular along z."""
    y = x.copy()
    y[m:] -= y[m:].mean()
    return y


def rel(a, b) -> float:
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def ortho_horizon(V: np.ndarray, thresh: float = 1e-8) -> int:
    """last step J with ||I - V_J^T V_J||_F < thresh (parity contract SURVEY 8c(i))."""
    g = V.T @ V
    for j in range(1, V.shape[1] + 1):
        if np.linalg.norm(np.eye(j) - g[:j, :j]) >= thresh:
            return j - 1
    return V.shape[1]


def stability_spectrum(n: int, func: str, scenario: str) -> np.ndarray:
    """diagonal spectra of src/bin/stability.rs:98-146 (identical in orthogonality.rs:91-146)."""
    i = np.arange(n, dtype=np.float64)
    d = float(max(n - 1, 1))
    if func == "exp" and scenario == "well":
        return -10.0 + (9.9 / d) * i
    if func == "exp" and scenario == "ill":
        return -1000.0 + (999.9 / d) * i
    if func == "inv" and scenario == "well":
        return 0.1 + (99.9 / d) * i
    if func == "inv" and scenario == "ill":
        mid = n // 2
        e = np.where(i < mid, 0.1 + (0.9 / max(mid - 1, 1)) * i, -1.0 + (0.9 / max(n - mid - 1, 1)) * (i - mid))
        e[mid] = 1e-8
        return e
    raise ValueError((func, scenario))


FTK = {"inv": npo.inv_tk_solver, "exp": npo.exp_tk_solver, "square": npo.square_tk_solver}
