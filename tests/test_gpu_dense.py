"""Dense symmetric operator (src/bin/dense_tradeoff.rs) through the C ABI against the CPU oracle run on the same
matrix as a full CSC: same parity contract as the sparse operators."""
import numpy as np
import pytest

import helpers
import two_pass_lanczos_b200 as tpl
from oracle import np_oracle as npo
from oracle import oracle as orc
from two_pass_lanczos_b200 import algorithms as alg

pytestmark = pytest.mark.gpu


def sym(n, seed, shift=0.0):
    g = np.random.default_rng(seed).standard_normal((n, n))
    return (g + g.T) / 2 + shift * np.eye(n)


def oracle_dense(a):
    n = a.shape[0]
    rows, cols = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(n, dtype=np.uint64), indexing="ij")
    return orc.SparseColMat.try_new_from_triplets(n, n, rows.ravel(), cols.ravel(), a.ravel())


@pytest.fixture(scope="module", params=[(97, 1), (1000, 2), (2500, 3)], ids=lambda p: f"n{p[0]}")
def case(request):
    n, seed = request.param
    a = sym(n, seed)
    return a, oracle_dense(a), tpl.LinOp.from_dense(a)


def test_format_and_apply(case):
    a, oop, gop = case
    assert gop.format == "dense" and gop.kernel_shape() == "dense" and gop.matrix_bytes() == 8 * a.shape[0] ** 2
    x = np.random.default_rng(5).standard_normal(a.shape[0])
    assert helpers.rel(gop.apply(x), a @ x) < 1e-13


def test_coefficients_basis_and_drift(case):
    a, oop, gop = case
    n = a.shape[0]
    k = min(60, n - 1)
    b = helpers.seeded_b(n)
    v_ref, d_ref = orc.lanczos_standard(oop, b, k)
    out = alg.lanczos_standard(gop, b, k)
    po = alg.lanczos_pass_one(gop, b, k)
    assert po.steps_taken == d_ref.steps_taken
    assert np.array_equal(po.alphas, out.decomposition.alphas) and np.array_equal(po.betas, out.decomposition.betas)
    J = helpers.ortho_horizon(v_ref, 1e-8)
    assert J >= 20
    assert np.max(np.abs(po.alphas[:J] - d_ref.alphas[:J])) <= 1e-12 * np.abs(d_ref.alphas).max()
    assert np.max(np.abs(po.betas[:J - 1] - d_ref.betas[:J - 1])) <= 1e-12 * np.abs(d_ref.betas).max()
    y = 0.1 * (np.arange(po.steps_taken) + 1)
    p2 = alg.lanczos_pass_two_with_basis(gop, b, po, y)
    assert np.array_equal(p2.v_k, out.v_k)          # regenerated basis bit-identical to the stored one
    assert helpers.rel(p2.x_k, out.v_k @ y) < 1e-13


def test_exp_two_pass_vs_oracle_and_one_pass(case):
    a, oop, gop = case
    n = a.shape[0]
    k = min(40, n - 1)
    a_s = a / np.abs(np.linalg.eigvalsh(a)).max() * 5.0  # moderate spectrum for exp
    gop_s, oop_s = tpl.LinOp.from_dense(a_s), oracle_dense(a_s)
    b = helpers.seeded_b(n)
    x2 = tpl.lanczos_two_pass(gop_s, b, k, "exp")
    x1 = tpl.lanczos(gop_s, b, k, "exp")
    x_ref = orc.lanczos_two_pass(oop_s, b, k, npo.exp_tk_solver)
    assert helpers.rel(x2, x_ref) < 1e-10
    assert helpers.rel(x1, x2) < 1e-12


def test_inv_spd_against_direct_solve():
    n, k = 1200, 120
    g = np.random.default_rng(9).standard_normal((n, n))
    a = g @ g.T / n + np.eye(n)                       # SPD, condition ~ 5
    gop = tpl.LinOp.from_dense(a)
    b = helpers.seeded_b(n)
    x = tpl.lanczos_two_pass(gop, b, k, "inv")
    assert helpers.rel(x, np.linalg.solve(a, b)) < 1e-10


def test_callback_and_breakdown():
    gop = tpl.LinOp.from_dense(np.diag([2.0, 3.0]))
    out = alg.lanczos_standard(gop, [1.0, 0.0], 2)
    assert out.decomposition.steps_taken == 1           # mod.rs:410-419
    a = sym(300, 4)
    gop = tpl.LinOp.from_dense(a)
    b = helpers.seeded_b(300)
    seen = []
    stopped = alg.lanczos_standard(gop, b, 20, callback=lambda k, v, t: (seen.append(k), k < 5)[1])
    full = alg.lanczos_standard(gop, b, 20)
    assert stopped.decomposition.steps_taken == 5 and seen == [1, 2, 3, 4, 5]
    assert np.array_equal(stopped.v_k, full.v_k[:, :5])
