"""Host f(T_k) e1 solvers of the library (tpl_ftk_*, C++) against dense numpy linear algebra, i.e. against what
the reference's closures compute with faer (tests/correctness.rs:171-299, src/bin/stability.rs:161-193)."""
import ctypes as C

import numpy as np
import pytest

from oracle import np_oracle as npo
from two_pass_lanczos_b200 import _lib


def call(name, alphas, betas):
    lib = _lib.load()
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    be = np.ascontiguousarray(betas, dtype=np.float64)
    y = np.zeros(max(len(al), 1))
    ylen = C.c_size_t(len(al))
    dp = _lib.c_dp
    al_ = al if len(al) else np.zeros(1)
    be_ = be if len(be) else np.zeros(1)
    rc = getattr(lib, name)(al_.ctypes.data_as(dp), len(al), be_.ctypes.data_as(dp), len(be), y.ctypes.data_as(dp),
                            C.byref(ylen), None)
    return rc, y[: ylen.value]


def lanczos_like_tridiag(k, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    return spread * rng.standard_normal(k), np.abs(rng.standard_normal(max(k - 1, 0))) + 0.05


@pytest.mark.parametrize("k", [1, 2, 3, 7, 30, 200, 500])
def test_inv(k):
    al, be = lanczos_like_tridiag(k, k)
    al = al + 4.0  # diagonally dominant-ish, well conditioned
    rc, y = call("tpl_ftk_inv", al, be)
    assert rc == 0 and len(y) == k
    ref = npo.inv_tk_solver(al, be)
    assert np.linalg.norm(y - ref) <= 1e-12 * np.linalg.norm(ref)


def test_inv_needs_pivoting():
    al = np.array([0.0, 0.0, 1.0, 0.0])
    be = np.array([1.0, 2.0, 3.0])
    rc, y = call("tpl_ftk_inv", al, be)
    assert rc == 0
    assert np.allclose(y, npo.inv_tk_solver(al, be), rtol=1e-13, atol=1e-15)
    rc, _ = call("tpl_ftk_inv", [0.0, 0.0], [0.0])  # singular T_k -> solver error
    assert rc != 0
    assert "singular" in _lib.load().tpl_last_error_message().decode()


@pytest.mark.parametrize("k", [1, 2, 5, 30, 200])
def test_exp(k):
    al, be = lanczos_like_tridiag(k, 100 + k)
    al = al - 3.0
    rc, y = call("tpl_ftk_exp", al, be)
    assert rc == 0 and len(y) == k
    ref = npo.exp_tk_solver(al, be)
    assert np.linalg.norm(y - ref) <= 1e-12 * np.linalg.norm(ref)


def test_exp_wide_spectrum():
    """T_150 of the ill-conditioned exp scenario (eigenvalues in [-1000, -0.1], src/bin/stability.rs:115-122)."""
    import scipy.sparse as sp

    import helpers

    n, k = 2000, 150
    eigs = helpers.stability_spectrum(n, "exp", "ill")
    al, be, steps, _ = npo.pass_one(sp.diags(eigs).tocsr(), helpers.seeded_b(n), k)
    assert steps == k
    rc, y = call("tpl_ftk_exp", al, be)
    assert rc == 0
    ref = npo.exp_tk_solver(al, be)
    assert np.linalg.norm(y - ref) <= 1e-10 * np.linalg.norm(ref)


@pytest.mark.parametrize("k", [1, 2, 3, 30])
def test_square(k):
    al, be = lanczos_like_tridiag(k, 200 + k)
    rc, y = call("tpl_ftk_square", al, be)
    assert rc == 0
    assert np.allclose(y, npo.square_tk_solver(al, be), rtol=1e-15, atol=0)


def test_empty():
    for name in ("tpl_ftk_inv", "tpl_ftk_exp", "tpl_ftk_square"):
        rc, y = call(name, [], [])
        assert rc == 0 and len(y) == 0  # `return Ok(Mat::zeros(0, 1))` in every reference closure


# ------------------------------------------------------------------ residual estimates
def _direct_residuals(al, be_full, b_norm):
    """||b|| beta_j |e_j^T T_j^{-1} e_1| by dense solves"""
    out = np.empty(len(al))
    for j in range(1, len(al) + 1):
        t = np.diag(al[:j]) + np.diag(be_full[: j - 1], 1) + np.diag(be_full[: j - 1], -1)
        e1 = np.zeros(j)
        e1[0] = 1.0
        out[j - 1] = b_norm * be_full[j - 1] * abs(np.linalg.solve(t, e1)[-1])
    return out


@pytest.mark.parametrize("k,shift", [(1, 4.0), (2, 4.0), (40, 4.0), (200, 4.0), (60, 0.0), (120, 0.3)])
def test_inv_residual_estimates_match_dense_solves(k, shift):
    """definite (shift 4) and indefinite (shift 0 / 0.3: some T_j nearly singular) tridiagonals"""
    from two_pass_lanczos_b200 import solvers

    al, be = lanczos_like_tridiag(k + 1, 100 + k)
    al, be = al[:k] + shift, be[:k]  # k betas: beta_k couples to v_{k+1}
    est = solvers.inv_residual_estimates(al, be, b_norm=3.0)
    ref = _direct_residuals(al, be, 3.0)
    assert est.shape == (k,) and np.all(np.isfinite(est))
    assert np.max(np.abs(est - ref) / ref) < 1e-8  # dense solves lose digits where T_j is nearly singular


def test_inv_residual_estimates_follow_a_real_lanczos_run():
    """on a well-conditioned diagonal operator the estimates equal the true residuals of the oracle's iterates"""
    from oracle import oracle as orc
    from two_pass_lanczos_b200 import solvers

    n, k = 400, 30
    eigs = np.linspace(1.0, 20.0, n)
    a = orc.SparseColMat.try_new_from_triplets(n, n, np.arange(n), np.arange(n), eigs)
    b = np.cos(np.arange(n)) + 2.0
    dec = orc.lanczos_pass_one(a, b, k + 1)
    est = solvers.inv_residual_estimates(dec.alphas[:k], dec.betas[:k], dec.b_norm)
    for j in (1, 5, 17, 30):
        xj = orc.lanczos_two_pass(a, b, j, npo.inv_tk_solver)
        true = np.linalg.norm(b - eigs * xj)
        assert abs(est[j - 1] - true) <= 1e-9 * np.linalg.norm(b) + 1e-6 * true, (j, est[j - 1], true)
    assert est[-1] < 1e-3 * est[0]  # and they decay


def test_inv_residual_estimates_edge_cases():
    from two_pass_lanczos_b200 import solvers

    assert solvers.inv_residual_estimates([], [], 1.0).shape == (0,)
    est = solvers.inv_residual_estimates([2.0, 3.0, 4.0], [1.0, 1.0], 1.0)  # reference-style: steps - 1 betas
    assert np.isfinite(est[0]) and np.isfinite(est[1]) and np.isnan(est[2])
    est = solvers.inv_residual_estimates([0.0, 1.0], [1.0, 1.0], 1.0)  # T_1 = [0] is singular: no first iterate
    assert np.isinf(est[0]) and np.isfinite(est[1])
    est = solvers.inv_residual_estimates([2.0, 3.0], [0.0, 1.0], 5.0)  # exact breakdown after one step: x_1 is exact
    assert est[0] == 0.0
