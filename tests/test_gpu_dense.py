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


# ---------------------------------------------------------------------------------------------------------------------
# Complex Hermitian operators (`T: ComplexField`, src/algorithms/mod.rs:167).  The reference's tests instantiate f64 only, so
# there is no reference-held vector for T = c64: parity is against a complex128 restatement of the SAME recurrence
# (mod.rs:292-340: w = A v - beta v_prev; alpha = Re <v, w>; w -= alpha v; beta = ||w||; v_next = w / beta) and of the two-pass
# solve (solvers.rs:133-175), i.e. "parity unpinned" for this operator kind.
def hermitian(n, seed):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    a = (g + g.conj().T) / np.sqrt(2.0 * n)
    a[np.diag_indices(n)] = a[np.diag_indices(n)].real
    return a


def complex_lanczos(a, b, k):
    n = len(b)
    V = np.zeros((n, k), dtype=np.complex128)
    al, be = np.zeros(k), np.zeros(max(k - 1, 0))
    bn = np.linalg.norm(b)
    v_prev, v, beta_prev = np.zeros(n, dtype=np.complex128), b / bn, 0.0
    for j in range(k):
        V[:, j] = v
        w = a @ v - beta_prev * v_prev
        al[j] = np.vdot(v, w).real
        w = w - al[j] * v
        beta = np.linalg.norm(w)
        if j + 1 < k:
            be[j] = beta
            v_prev, v, beta_prev = v, w / beta, beta
    return V, al, be, bn


@pytest.mark.parametrize("n", [5, 257, 1500, 13001])
def test_hermitian_operator_against_complex_restatement(n):
    a = hermitian(n, n)
    gop = tpl.LinOp.from_dense_hermitian(a)
    assert gop.is_complex and gop.nrows() == 2 * n and gop.kernel_shape() == "dense"
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = gop.apply(x)
    assert y.dtype == np.complex128 and helpers.rel(y, a @ x) < 1e-14
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    k = min(30, n - 1)
    V, al, be, bn = complex_lanczos(a, b, k)
    dec = alg.lanczos_pass_one(gop, b, k)
    assert dec.steps_taken == k and abs(dec.b_norm - bn) <= 1e-14 * bn
    assert np.max(np.abs(dec.alphas - al)) <= 1e-12 * np.abs(al).max()
    assert np.max(np.abs(dec.betas - be)) <= 1e-12 * np.abs(be).max()
    # f(A) b = ||b|| V f(T_k) e1 with the real T_k of the complex recurrence (solvers.rs:46-107, 133-175)
    T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
    lam, q = np.linalg.eigh(T)
    for f, fun in (("exp", np.exp), ("inv", lambda z: 1.0 / z)):
        if f == "inv":
            shift = 3.0 * np.eye(n)  # Hermitian positive definite
            gop_f, a_f = tpl.LinOp.from_dense_hermitian(a + shift), a + shift
            Vf, alf, bef, bnf = complex_lanczos(a_f, b, k)
            lamf, qf = np.linalg.eigh(np.diag(alf) + np.diag(bef, 1) + np.diag(bef, -1))
            want = bnf * (Vf @ (qf @ (fun(lamf) * qf[0, :])))
        else:
            gop_f, want = gop, bn * (V @ (q @ (fun(lam) * q[0, :])))
        x2 = tpl.lanczos_two_pass(gop_f, b, k, f)
        x1 = tpl.lanczos(gop_f, b, k, f)
        assert x2.dtype == np.complex128 and x2.shape == (n,)
        assert helpers.rel(x2, want) < 1e-10
        assert helpers.rel(x1, x2) < 1e-12
    if n <= 1500:  # converged Krylov approximation against the dense function of the matrix
        lam_a, q_a = np.linalg.eigh(a)
        xk = tpl.lanczos_two_pass(gop, b, min(n, 60), "exp")
        assert helpers.rel(xk, q_a @ (np.exp(lam_a) * (q_a.conj().T @ b))) < 1e-9


def test_hermitian_real_case_equals_the_symmetric_operator():
    """A Hermitian operator with zero imaginary parts and a real right-hand side gives the symmetric operator's iterates."""
    a = sym(700, 3)
    b = helpers.seeded_b(700)
    xr = tpl.lanczos_two_pass(tpl.LinOp.from_dense(a), b, 40, "exp")
    xc = tpl.lanczos_two_pass(tpl.LinOp.from_dense_hermitian(a.astype(np.complex128)), b, 40, "exp")
    assert np.max(np.abs(xc.imag)) == 0.0 and helpers.rel(xc.real, xr) < 1e-12


def test_diagonal_operator():
    """diag(d) behind the same boundary (the spectra of src/bin/stability.rs:98-193)."""
    n = 20000
    lam = helpers.stability_spectrum(n, "exp", "well")
    gop = tpl.LinOp.from_diagonal(lam)
    b = helpers.reference_b(n)
    assert np.array_equal(gop.apply(b), lam * b)
    x = tpl.lanczos_two_pass(gop, b, 60, "exp")
    assert helpers.rel(x, np.exp(lam) * b) < 1e-9
