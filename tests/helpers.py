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
    """b of the reference's tests and benches: `StdRng::seed_from_u64(42)` uniforms (two_pass_lanczos_b200/stdrng.py; pinned
    end to end by the reference's own published outputs, tests/test_oracle_reference_kats.py::test_published_accuracy_rows)."""
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
    from two_pass_lanczos_b200 import experiments

    return experiments.diagonal_spectrum(n, func, scenario)


def published_curves() -> dict:
    """Numbers of the reference's results/accuracy_*.csv and results/orthogonality_*.csv (outputs of src/bin/stability.rs and
    src/bin/orthogonality.rs), committed as tests/golden/published_curves.json by tests/golden/make_published_curves.py."""
    import json

    with open(os.path.join(GOLDEN, "published_curves.json")) as fh:
        return json.load(fh)


# Relative tolerance on a published error value, per curve: how closely a faithful implementation reproduces the reference's
# own printed relative errors (rows with a published error above ACCURACY_FLOOR; below it both sides sit on rounding noise).
# Measured for the oracle: 1.1e-10 / 2.1e-13 / 2.7e-4 / 6.4e-5 (the ill-conditioned curves amplify the last-bit differences
# of the dot-product order by 1e9 ... 1e12 at their tails).
ACCURACY_RTOL = {"inv_well": 1e-8, "exp_well": 1e-10, "exp_ill": 1e-3, "inv_ill": 1e-3}
ACCURACY_FLOOR = 1e-11


def check_accuracy_row(curve: str, k: int, published: float, measured: float) -> None:
    if published > ACCURACY_FLOOR:
        assert abs(measured - published) <= ACCURACY_RTOL[curve] * published, (curve, k, published, measured)
    else:
        assert measured <= 20.0 * ACCURACY_FLOOR, (curve, k, published, measured)


def cpu_spread(inst: datagen.KKTInstance, b: np.ndarray, k: int, x_cpu: np.ndarray, ftk=None) -> float:
    """How far two LEGITIMATE CPU evaluations of the same reference path differ on this problem: the numpy restatement (CSR row
    sums, pairwise dot products) against the C++ oracle's x.  On converged, well-conditioned problems this is ~1e-15; on an
    ill-conditioned or unconverged one (50 k arcs 'aa' at k = 500: 2e-5) it is the sensitivity of x to the last bit of the
    summation order, which no implementation -- faer included -- can be asked to beat."""
    a_sp = npo.kkt_matrix(inst.m, inst.p, inst.tail.astype(np.int64), inst.head.astype(np.int64), inst.d)
    return rel(npo.lanczos_two_pass(a_sp, b, k, ftk or npo.inv_tk_solver), x_cpu)


FTK = {"inv": npo.inv_tk_solver, "exp": npo.exp_tk_solver, "square": npo.square_tk_solver}
