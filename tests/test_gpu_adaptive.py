"""residual estimates of all Lanczos iterates from one pass 1, and the two-pass solve whose k is chosen from
them (tpl_ftk_inv_residuals, tpl_lanczos_two_pass_inv_adaptive).  No reference counterpart; checked against true residuals
formed with the operator and against the fixed-k solver."""
import numpy as np
import pytest

import helpers
import two_pass_lanczos_b200 as tpl
from two_pass_lanczos_b200 import algorithms as alg
from two_pass_lanczos_b200 import datagen, solvers

pytestmark = pytest.mark.gpu


def _true_rel_residual(op, x, b):
    return np.linalg.norm(op.apply(x) - b) / np.linalg.norm(b)


@pytest.mark.parametrize("m", [5_000, 500_000])
def test_estimates_track_true_residuals_on_a_kkt_instance(m):
    inst = datagen.gen_kkt(m, 3, 2, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    k = 120
    dec = alg.lanczos_pass_one(op, b, k + 1)
    est = solvers.inv_residual_estimates(dec.alphas[:k], dec.betas[:k], dec.b_norm) / np.linalg.norm(b)
    for j in (5, 20, 60, 120):
        true = _true_rel_residual(op, tpl.lanczos_two_pass(op, b, j, "inv"), b)
        # equal while the basis is orthogonal; afterwards the estimate may run ahead of the true residual's plateau
        assert true <= 3.0 * est[j - 1] + 1e-9, (j, true, est[j - 1])
        if est[j - 1] > 1e-7:
            assert est[j - 1] <= 3.0 * true, (j, true, est[j - 1])


def test_adaptive_solve_stops_early_and_equals_the_fixed_k_solve():
    inst = datagen.gen_kkt(20_000, 3, 4, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    x, k_used, est = solvers.lanczos_two_pass_inv_adaptive(op, b, 400, 1e-6)
    assert 1 <= k_used < 400
    assert est <= 1e-6 * np.linalg.norm(b)
    assert _true_rel_residual(op, x, b) <= 1e-5
    assert np.array_equal(x, tpl.lanczos_two_pass(op, b, k_used, "inv"))  # same coefficients, same pass 2
    # one step earlier the estimate was still above the tolerance
    dec = alg.lanczos_pass_one(op, b, k_used + 1)
    all_est = solvers.inv_residual_estimates(dec.alphas[:k_used], dec.betas[:k_used], dec.b_norm)
    assert np.all(all_est[: k_used - 1] > 1e-6 * np.linalg.norm(b))
    assert all_est[k_used - 1] == pytest.approx(est, rel=1e-12)


def test_adaptive_solve_without_convergence_returns_the_best_iterate():
    inst = datagen.gen_kkt(20_000, 3, 4, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    k_max = 12
    x, k_used, est = solvers.lanczos_two_pass_inv_adaptive(op, b, k_max, 0.0)
    assert 1 <= k_used <= k_max
    dec = alg.lanczos_pass_one(op, b, k_max + 1)
    all_est = solvers.inv_residual_estimates(dec.alphas[:k_max], dec.betas[:k_max], dec.b_norm)
    assert k_used == int(np.argmin(all_est)) + 1 and est == pytest.approx(all_est.min(), rel=1e-12)
    assert np.array_equal(x, tpl.lanczos_two_pass(op, b, k_used, "inv"))


def test_adaptive_solve_device_vectors_and_errors():
    import torch

    inst = datagen.gen_kkt(5_000, 3, 1, "wc")
    op = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    x, k_used, _ = solvers.lanczos_two_pass_inv_adaptive(op, b, 200, 1e-6)
    xd, kd, _ = solvers.lanczos_two_pass_inv_adaptive(op, torch.from_numpy(b).cuda(), 200, 1e-6)
    assert kd == k_used and np.array_equal(xd.cpu().numpy(), x)
    with pytest.raises(tpl.LanczosError):
        solvers.lanczos_two_pass_inv_adaptive(op, b, 200, float("nan"))
    with pytest.raises(tpl.LanczosError):
        solvers.lanczos_two_pass_inv_adaptive(op, b, 0, 1e-6)  # k == 0: the reference panics
    with pytest.raises(tpl.LanczosError) as e:  # zero b is an InputError, as in pass 1 of the reference (mod.rs:229-235)
        solvers.lanczos_two_pass_inv_adaptive(op, np.zeros(inst.n), 10, 1e-6)
    assert e.value.kind == "InputError"
