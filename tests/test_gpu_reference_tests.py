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


def test_callback_sees_the_breakdown_step():
    """lanczos.rs:93-112: the callback runs after EVERY completed step, also the one whose beta is <= tol; the breakdown is
    looked at afterwards.  diag(2, 3), b = e1 breaks down in step 1 (mod.rs:410-419)."""
    a = tpl.LinOp.from_dense(np.diag([2.0, 3.0]))
    seen = []
    out = alg.lanczos_standard(a, [1, 0], 2, callback=lambda k, v_k, t_k: seen.append((k, len(t_k.alphas), len(t_k.betas))) or True)
    assert out.decomposition.steps_taken == 1
    assert seen == [(1, 1, 0)]
    with pytest.raises(tpl.LanczosError):  # a zero b completes no step: no callback
        alg.lanczos_standard(a, [0, 0], 2, callback=lambda *args: seen.append("never") or True)
    assert seen == [(1, 1, 0)]


def test_dimension_mismatch_is_an_error_not_an_overrun():
    """The C entry points take bare pointers; the binding checks every vector against nrows() first and reports the
    reference's DimensionMismatch (src/error.rs:29-35; faer panics on the same misuse)."""
    a, b, _ = simple_problem()
    for fn in (lambda: tpl.lanczos_two_pass(a, b[:3], 3, "inv"), lambda: tpl.lanczos(a, np.ones(5), 3, "inv"),
               lambda: alg.lanczos_pass_one(a, b[:2], 3), lambda: alg.lanczos_standard(a, np.ones(9), 3),
               lambda: a.apply(np.ones(3))):
        with pytest.raises(tpl.LanczosError) as e:
            fn()
        assert e.value.kind == "DimensionMismatch"
    with pytest.raises(tpl.LanczosError) as e:
        a.apply(np.ones(3))
    assert str(e.value) == "Dimension mismatch: operator has 4 columns but vector has 3 rows."
    d = alg.lanczos_pass_one(a, b, 3)
    short = alg.LanczosDecomposition(d.alphas[:2], d.betas, 3, d.b_norm)
    with pytest.raises(tpl.LanczosError) as e:
        alg.lanczos_pass_two(a, b, short, np.ones(3))
    assert e.value.kind == "ParameterMismatch"
    short = alg.LanczosDecomposition(d.alphas, d.betas[:1], 3, d.b_norm)
    with pytest.raises(tpl.LanczosError):
        alg.lanczos_pass_two(a, b, short, np.ones(3))


def test_handles_are_independent_in_shared_memory_size():
    """A later handle with a small shared-memory footprint must not lower the opt-in cap of the kernels under an earlier
    handle that needs more (the attribute belongs to the kernel, not to the handle): alternate solves on a large-p and a
    small-p operator of every incidence shape."""
    from two_pass_lanczos_b200 import datagen

    big = datagen.gen_kkt(2_000_000, 3, 3, "wc")      # p = 2309: large node segment, tiled kernels with big tiles
    small = datagen.gen_kkt(3_000, 3, 4, "wc")        # p = 89
    op_big = tpl.LinOp.from_kkt(big.m, big.p, big.tail, big.head, big.d)
    b_big = op_big.apply(np.full(big.n, 1.0 / np.sqrt(big.n)))
    ref = {}
    for mode in (0, 2, 3):
        op_big.set_mode(mode)
        ref[mode] = alg.lanczos_pass_one(op_big, b_big, 12)
    op_small = tpl.LinOp.from_kkt(small.m, small.p, small.tail, small.head, small.d)  # created AFTER the large one
    b_small = op_small.apply(np.full(small.n, 1.0 / np.sqrt(small.n)))
    for mode in (0, 2, 3, 4):
        op_small.set_mode(mode)
        alg.lanczos_pass_one(op_small, b_small, 12)
        x_small = tpl.lanczos_two_pass(op_small, b_small, 12, "exp")
        assert np.all(np.isfinite(x_small))
    for mode in (0, 2, 3):  # the large handle still launches, with the same results
        op_big.set_mode(mode)
        again = alg.lanczos_pass_one(op_big, b_big, 12)
        assert np.array_equal(again.alphas, ref[mode].alphas) and np.array_equal(again.betas, ref[mode].betas)
    op_big.close()
    op_small.close()


def test_cuda_tensor_inputs_are_ordered_after_their_producer():
    """A CUDA tensor produced on torch's current stream just before the call is consumed correctly without any explicit
    synchronisation: the operator adopts the current stream (operators.LinOp._vec)."""
    import torch

    from two_pass_lanczos_b200 import datagen

    inst = datagen.gen_kkt(200_000, 3, 5, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    x0 = np.full(inst.n, 1.0 / np.sqrt(inst.n))
    b_host = op.apply(x0)
    expect = tpl.lanczos_two_pass(op, b_host, 25, "exp")
    side = torch.cuda.Stream()
    for stream in (torch.cuda.current_stream(), side):
        with torch.cuda.stream(stream):
            big = torch.ones(64 << 20, dtype=torch.float64, device="cuda")
            for _ in range(4):          # a few ms of queued work in front of the producer of b
                big.mul_(1.0000001)
            xt = torch.full((inst.n,), 1.0 / np.sqrt(inst.n), dtype=torch.float64, device="cuda")
            bt = op.apply(xt) * 1.0     # produced on `stream`, consumed by the library right away
            x = tpl.lanczos_two_pass(op, bt, 25, "exp")
            got = x.cpu().numpy()
        assert np.array_equal(got, expect)
    op.close()
