"""The reference's own tests, re-expressed through the C ABI on the GPU (names follow the Rust tests):
src/algorithms/mod.rs:384-428, tests/correctness.rs:165-325, src/lib.rs:35-84, plus the error behaviour of
src/solvers.rs and src/algorithms/lanczos_two_pass.rs:220-244."""
import numpy as np
import pytest

import helpers
import two_pass_lanczos_b200 as tpl
from two_pass_lanczos_b200 import algorithms as alg

pytestmark = pytest.mark.gpu

APPROX_TOLERANCE = 1e-3   # tests/correctness.rs:42
EXACT_TOLERANCE = 1e-12   # tests/correctness.rs:51


def simple_problem():
    a = np.array([[2, -1, 0, 0], [-1, 2, -1, 0], [0, -1, 2, -1], [0, 0, -1, 2.0]])
    return tpl.LinOp.from_dense(a), np.array([1, 2, 3, 4.0]), a


def test_recurrence_step_correctness():  # mod.rs:385-407
    a, _, _ = simple_problem()
    d = alg.lanczos_pass_one(a, [1, 0, 0, 0], 2)
    assert abs(d.alphas[0] - 2.0) < 1e-15
    assert abs(d.betas[0] - 1.0) < 1e-15


def test_breakdown_scenario():  # mod.rs:410-419
    a = tpl.LinOp.from_dense(np.diag([2.0, 3.0]))
    out = alg.lanczos_standard(a, [1, 0], 2)
    assert out.decomposition.steps_taken == 1
    assert out.v_k.shape == (2, 1) and len(out.decomposition.betas) == 0
    po = alg.lanczos_pass_one(a, [1, 0], 2)
    assert po.steps_taken == 1 and len(po.alphas) == 1 and len(po.betas) == 0
    # two-pass on a broken-down decomposition still works: x = y_0 v_1 * ... (steps == 1 -> no regeneration)
    x = tpl.lanczos_two_pass(a, [1, 0], 2, "inv")
    assert np.allclose(x, [0.5, 0.0], rtol=1e-15)


def test_zero_vector_input_returns_error():  # mod.rs:422-428
    a = tpl.LinOp.from_dense(np.eye(2))
    for fn in (lambda: alg.lanczos_standard(a, [0, 0], 2), lambda: alg.lanczos_pass_one(a, [0, 0], 2),
               lambda: tpl.lanczos(a, [0, 0], 2, "inv"), lambda: tpl.lanczos_two_pass(a, [0, 0], 2, "inv")):
        with pytest.raises(tpl.LanczosError) as e:
            fn()
        assert e.value.kind == "InputError"
        assert str(e.value) == "Invalid input parameter: Input vector `b` must not be a zero vector."


def test_doctest():  # src/lib.rs:35-84
    a, b, dense = simple_problem()
    x1 = tpl.lanczos(a, b, 3, helpers.FTK["inv"])
    x2 = tpl.lanczos_two_pass(a, b, 3, helpers.FTK["inv"])
    assert np.linalg.norm(x1 - x2) < 1e-12
    x4 = tpl.lanczos_two_pass(a, b, 4, "inv")  # k = n: exact solve
    assert np.allclose(x4, np.linalg.solve(dense, b), rtol=1e-12)


@pytest.mark.parametrize("solver", ["lanczos", "lanczos_two_pass"])
@pytest.mark.parametrize("fname,f,tol", [("inv", lambda z: 1.0 / z, APPROX_TOLERANCE),
                                         ("exp", np.exp, APPROX_TOLERANCE),
                                         ("square", lambda z: z * z, EXACT_TOLERANCE)])
@pytest.mark.parametrize("closure,bgen", [("python", "numpy"), ("native", "numpy"), ("native", "stdrng")])
def test_correctness_rs(solver, fname, f, tol, closure, bgen):  # tests/correctness.rs:165-325
    n, k = 100, 30
    eigs = np.arange(1, n + 1.0)
    import scipy.sparse as sp

    a = tpl.LinOp.from_scipy(sp.diags(eigs))
    b = helpers.B_GENERATORS[bgen](n)
    x_true = f(eigs) * b
    ftk = helpers.FTK[fname] if closure == "python" else fname
    x = getattr(tpl, solver)(a, b, k, ftk)
    assert helpers.rel(x, x_true) < tol


def test_error_behaviour():
    a, b, _ = simple_problem()
    d = alg.lanczos_pass_one(a, b, 3)
    with pytest.raises(tpl.LanczosError) as e:  # lanczos_two_pass.rs:220-227
        alg.lanczos_pass_two(a, b, d, np.ones(2))
    assert e.value.kind == "ParameterMismatch"
    assert str(e.value) == "Parameter mismatch: `y_k` expects size 3, but got 2."
    d0 = alg.LanczosDecomposition(d.alphas, d.betas, d.steps_taken, 0.0)
    with pytest.raises(tpl.LanczosError) as e:  # lanczos_two_pass.rs:229-235
        alg.lanczos_pass_two(a, b, d0, np.ones(3))
    assert str(e.value) == "Invalid input parameter: The initial vector `b` must not be a zero vector."
    dz = alg.LanczosDecomposition(np.zeros(0), np.zeros(0), 0, d.b_norm)
    assert np.array_equal(alg.lanczos_pass_two(a, b, dz, np.zeros(0)), np.zeros(4))  # :237-244
    for solver in (tpl.lanczos, tpl.lanczos_two_pass):
        with pytest.raises(tpl.LanczosError) as e:  # solvers.rs:75-82, 158-165
            solver(a, b, 3, lambda al, be: np.ones(5))
        assert str(e.value) == "Parameter mismatch: `y_k_prime` expects size 3, but got 5."
        with pytest.raises(tpl.LanczosError) as e:  # solvers.rs:72, 156
            solver(a, b, 3, lambda al, be: (_ for _ in ()).throw(RuntimeError("Custom solver failed")))
        assert str(e.value) == "The user-provided f(T_k) solver failed: Custom solver failed"  # error.rs:120-128
        with pytest.raises(tpl.LanczosError) as e:  # k == 0: the reference panics
            solver(a, b, 0, "inv")
        assert e.value.kind == "Panic"


def test_callback_early_stop():  # LanczosCallback, lanczos.rs:93-106
    n = 200
    import scipy.sparse as sp

    a = tpl.LinOp.from_scipy(sp.diags(np.arange(1, n + 1.0)))
    b = helpers.seeded_b(n)
    seen = []

    def cb(k, v_k, t_k):
        seen.append((k, len(t_k.alphas), len(t_k.betas), v_k.steps))
        return k < 7

    out = alg.lanczos_standard(a, b, 20, callback=cb)
    assert out.decomposition.steps_taken == 7
    assert len(out.decomposition.alphas) == 7 and len(out.decomposition.betas) == 6
    assert seen == [(i, i, i - 1, i) for i in range(1, 8)]
    full = alg.lanczos_standard(a, b, 20)
    assert np.array_equal(out.v_k, full.v_k[:, :7])  # per-step launches reproduce the persistent kernel bitwise
    assert np.array_equal(out.decomposition.alphas, full.decomposition.alphas[:7])
