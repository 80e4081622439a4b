"""The reference's accuracy and orthogonality experiments (src/bin/stability.rs, src/bin/orthogonality.rs) on the GPU engine:
same synthetic diagonal problems, same right-hand side (`StdRng::seed_from_u64(42)` uniforms), same metrics, same CSV schemas
(results/accuracy_*.csv, results/orthogonality_*.csv).  `scripts/stability.py` and `scripts/orthogonality.py` are the CLIs."""
from __future__ import annotations

import numpy as np

from . import algorithms as alg
from . import solvers, stdrng
from .operators import LinOp

FUNCTIONS = ("inv", "exp")
SCENARIOS = ("well-conditioned", "ill-conditioned")
ACCURACY_COLUMNS = ("k", "relative_error_standard", "relative_error_two_pass", "relative_solution_deviation")
ORTHOGONALITY_COLUMNS = ("k", "ortho_loss_standard", "ortho_loss_regenerated", "basis_drift_fro", "solution_deviation_l2")


def diagonal_spectrum(n: int, function: str, scenario: str) -> np.ndarray:
    """`create_diagonal_problem` (src/bin/stability.rs:98-146, identical in src/bin/orthogonality.rs:91-146)."""
    scenario = {"well": "well-conditioned", "ill": "ill-conditioned"}.get(scenario, scenario)
    if function not in FUNCTIONS or scenario not in SCENARIOS:
        raise ValueError((function, scenario))
    i = np.arange(n, dtype=np.float64)
    d = float(max(n - 1, 1))
    if function == "exp":
        return (-10.0 + (9.9 / d) * i) if scenario == "well-conditioned" else (-1000.0 + (999.9 / d) * i)
    if scenario == "well-conditioned":
        return 0.1 + (99.9 / d) * i
    mid = n // 2
    e = np.where(i < mid, 0.1 + (0.9 / max(mid - 1, 1)) * i, -1.0 + (0.9 / max(n - mid - 1, 1)) * (i - mid))
    e[mid] = 1e-8  # "the critical eigenvalue"
    return e


def diagonal_operator(eigs: np.ndarray, device: int = -1) -> LinOp:
    """The sparse diagonal matrix of the experiments as a (generic CSR) device operator."""
    n = len(eigs)
    idx = np.arange(n + 1, dtype=np.uint64)
    return LinOp.from_csc(n, idx, idx[:-1], np.ascontiguousarray(eigs, dtype=np.float64), device)


def reference_rhs(n: int, seed: int = 42) -> np.ndarray:
    """`let mut rng = StdRng::seed_from_u64(42); Mat::from_fn(n, 1, |_, _| rng.random())` (stability.rs:256-257)."""
    return stdrng.std_rng_uniform(seed, n)


def _norm(v) -> float:
    return float(np.linalg.norm(v))


def run_accuracy(function: str, scenario: str, n: int = 1000, k_min: int = 5, k_max: int = 200, k_step: int = 5,
                 device: int = -1, f_tk_solver=None):
    """src/bin/stability.rs:244-323: for k = k_min, k_min + k_step, ... <= k_max the relative error of the one-pass and of the
    two-pass solution against the analytic f(lambda_i) b_i and their mutual deviation.  Rows follow ACCURACY_COLUMNS.  A k for
    which a solver fails is skipped, as in the reference (stability.rs:279-297)."""
    from .error import LanczosError

    eigs = diagonal_spectrum(n, function, scenario)
    op = diagonal_operator(eigs, device)
    b = reference_rhs(n)
    x_true = (np.exp(eigs) if function == "exp" else 1.0 / eigs) * b
    x_true_norm = _norm(x_true)
    f = f_tk_solver if f_tk_solver is not None else function
    rows = []
    try:
        for k in range(k_min, k_max + 1, k_step):
            if k == 0:
                continue
            try:
                x1 = np.asarray(solvers.lanczos(op, b, k, f))
                x2 = np.asarray(solvers.lanczos_two_pass(op, b, k, f))
            except LanczosError:
                continue
            rows.append((k, _norm(x1 - x_true) / x_true_norm, _norm(x2 - x_true) / x_true_norm, _norm(x1 - x2) / _norm(x1)))
    finally:
        op.close()
    return rows


def run_orthogonality(function: str, scenario: str, n: int = 1000, k_min: int = 20, k_max: int = 500, k_step: int = 20,
                      device: int = -1):
    """src/bin/orthogonality.rs:148-232: ||I - V^T V||_F of the stored basis (lanczos_standard) and of the regenerated one
    (lanczos_pass_two_with_basis with a dummy y = 0), their Frobenius distance and the distance of the two (zero) solutions.
    Rows follow ORTHOGONALITY_COLUMNS; the k column is steps_taken (orthogonality.rs:215)."""
    op = diagonal_operator(diagonal_spectrum(n, function, scenario), device)
    b = reference_rhs(n)
    rows = []
    try:
        for k in range(k_min, k_max + 1, k_step):
            if k == 0:
                continue
            out = alg.lanczos_standard(op, b, k)
            steps = out.decomposition.steps_taken
            if steps == 0:
                continue
            y = np.zeros(steps)
            p2 = alg.lanczos_pass_two_with_basis(op, b, out.decomposition, y)
            eye = np.eye(steps)
            rows.append((steps, _norm(eye - out.v_k.T @ out.v_k), _norm(eye - p2.v_k.T @ p2.v_k), _norm(out.v_k - p2.v_k),
                         _norm(out.v_k @ y - p2.v_k @ y)))
    finally:
        op.close()
    return rows


def write_csv(path: str, columns, rows) -> None:
    with open(path, "w") as fh:
        fh.write(",".join(columns) + "\n")
        for r in rows:
            fh.write(",".join(str(int(v)) if i == 0 else repr(float(v)) for i, v in enumerate(r)) + "\n")
