"""Second, independent restatement of the reference path in numpy/scipy (vectorised leaf ops).

TEST INFRASTRUCTURE ONLY (same rules as oracle/oracle.py).  Used to cross-check the C++ oracle with a
different summation order (pairwise numpy reductions, CSR row sums) -- the two oracles bracket the
"any deterministic order" freedom that faer's un-vendored kernels leave open --
and to supply the host f(T_k)e1 closures the reference's tests and benches use.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

TOL = 1000.0 * np.finfo(np.float64).eps  # breakdown_tolerance, src/algorithms/mod.rs:140-143


# ---- f(T_k) e1 closures ------------------------------------------------------------------------
def assemble_tridiagonal(alphas, betas) -> np.ndarray:  # tests/correctness.rs:65-79
    k = len(alphas)
    t = np.zeros((k, k))
    t[np.arange(k), np.arange(k)] = alphas
    if k > 1:
        i = np.arange(len(betas))
        t[i, i + 1] = betas
        t[i + 1, i] = betas
    return t


def inv_tk_solver(alphas, betas) -> np.ndarray:
    """T_k y = e1 by dense partial-pivot LU (tests/correctness.rs:171-179; stability.rs:161-170 uses sp_lu)."""
    k = len(alphas)
    if k == 0:
        return np.zeros(0)
    e1 = np.zeros(k)
    e1[0] = 1.0
    return np.linalg.solve(assemble_tridiagonal(alphas, betas), e1)


def exp_tk_solver(alphas, betas) -> np.ndarray:
    """Q exp(L) Q^T e1 via symmetric EVD (tests/correctness.rs:214-241, stability.rs:175-193)."""
    k = len(alphas)
    if k == 0:
        return np.zeros(0)
    lam, q = np.linalg.eigh(assemble_tridiagonal(alphas, betas))
    return q @ (np.exp(lam) * q[0, :])


def square_tk_solver(alphas, betas) -> np.ndarray:
    """T*T*e1 (tests/correctness.rs:290-299)."""
    k = len(alphas)
    if k == 0:
        return np.zeros(0)
    t = assemble_tridiagonal(alphas, betas)
    e1 = np.zeros(k)
    e1[0] = 1.0
    return (t @ t) @ e1


# ---- Lanczos (SURVEY Appendix A recipe) ------------------------------------------------------------
def pass_one(a: sp.spmatrix, b: np.ndarray, k: int):
    """src/algorithms/lanczos_two_pass.rs:65-110 -> (alphas, betas, steps, b_norm)."""
    if k == 0:
        raise OverflowError("capacity overflow")
    a = sp.csr_matrix(a)
    bn = float(np.linalg.norm(b))
    if bn <= TOL:
        raise ValueError("Invalid input parameter: Input vector `b` must not be a zero vector.")
    v = b * (1.0 / bn)
    vp = np.zeros_like(v)
    bp = 0.0
    alphas, betas = [], []
    steps = 0
    for i in range(k):
        w = a @ v
        w = w - bp * vp
        al = float(v @ w)
        w = w - al * v
        be = float(np.linalg.norm(w))
        alphas.append(al)
        steps += 1
        if be <= TOL:
            break
        w = w * (1.0 / be)
        vp, v, bp = v, w, be
        if i < k - 1:
            betas.append(be)
    return np.array(alphas), np.array(betas), steps, bn


def pass_two(a, b, alphas, betas, steps, bn, y, with_basis=False):
    """src/algorithms/lanczos_two_pass.rs:206-312."""
    a = sp.csr_matrix(a)
    if steps != len(y):
        raise ValueError(f"Parameter mismatch: `y_k` expects size {steps}, but got {len(y)}.")
    if bn <= TOL:
        raise ValueError("Invalid input parameter: The initial vector `b` must not be a zero vector.")
    n = len(b)
    if steps == 0:
        return (np.zeros(n), np.zeros((n, 0))) if with_basis else np.zeros(n)
    v = b * (1.0 / bn)
    vp = np.zeros_like(v)
    x = v * y[0]
    cols = [v]
    for j in range(steps - 1):
        bp = 0.0 if j == 0 else betas[j - 1]
        w = a @ v
        w = w - bp * vp
        w = w - alphas[j] * v
        w = w * (1.0 / betas[j])
        x = x + y[j + 1] * w
        vp, v = v, w
        if with_basis:
            cols.append(v)
    return (x, np.stack(cols, axis=1)) if with_basis else x


def lanczos_two_pass(a, b, k, f_tk_solver):
    """src/solvers.rs:133-175."""
    al, be, steps, bn = pass_one(a, b, k)
    if steps == 0:
        return np.zeros(len(b))
    yp = np.asarray(f_tk_solver(al, be)).reshape(-1)
    if len(yp) != steps:
        raise ValueError(f"Parameter mismatch: `y_k_prime` expects size {steps}, but got {len(yp)}.")
    return pass_two(a, b, al, be, steps, bn, yp * bn)


def kkt_matrix(m, p, tail, head, d) -> sp.csr_matrix:
    """A = [[D, E^T], [E, 0]] with the reference's index map."""
    j = np.arange(m)
    rows = np.concatenate([j[: len(d)], m + tail, m + head, j, j])
    cols = np.concatenate([j[: len(d)], j, j, m + tail, m + head])
    ones = np.ones(m)
    vals = np.concatenate([np.asarray(d, float), ones, -ones, ones, -ones])
    return sp.csr_matrix((vals, (rows, cols)), shape=(m + p, m + p))
