"""k-sweeps: every x_k of `lanczos_sweep` / `lanczos_two_pass_sweep` is bit-identical to the corresponding single
solve -- the reference's benches re-solve for every k (src/bin/tradeoff.rs:262-290, src/bin/stability.rs:259-312)."""
import numpy as np
import pytest

import helpers
import two_pass_lanczos_b200 as tpl
from two_pass_lanczos_b200 import datagen

pytestmark = pytest.mark.gpu


def _ops():
    inst = datagen.gen_kkt(50_000, 3, 3, "wc")
    cp, ri, va = datagen.kkt_csc(inst)
    ops = {"cells": tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d), "csr": tpl.LinOp.from_csc(inst.n, cp, ri, va)}
    blocked = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    blocked.set_mode(5)
    ops["blocked"] = blocked
    return inst, ops


@pytest.mark.parametrize("shape", ["cells", "blocked", "csr"])
def test_sweeps_are_bit_identical_to_single_solves(shape):
    inst, ops = _ops()
    op = ops[shape]
    b = op.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))
    ks = [5, 40, 41, 90, 200] + list(range(100, 118))  # 23 values: two chunks of the one-pass reconstruction kernel
    launches0 = op.kernel_launches()
    X1 = tpl.lanczos_sweep(op, b, ks, "inv")
    one_pass_launches = op.kernel_launches() - launches0
    X2 = tpl.lanczos_two_pass_sweep(op, b, ks, "inv")
    assert X1.shape == X2.shape == (len(ks), inst.n)
    assert one_pass_launches <= 4  # one basis generation + two reconstruction launches, not len(ks) solves
    for q, k in enumerate(ks):
        assert np.array_equal(X1[q], tpl.lanczos(op, b, k, "inv")), (shape, k)
        assert np.array_equal(X2[q], tpl.lanczos_two_pass(op, b, k, "inv")), (shape, k)
    assert helpers.rel(X1[4], X2[4]) < 1e-12
    for o in ops.values():
        o.close()


def test_sweep_closures_device_output_breakdown_and_errors():
    import scipy.sparse as sp
    import torch

    n = 300
    eigs = np.arange(1, n + 1.0)
    op = tpl.LinOp.from_scipy(sp.diags(eigs))
    b = helpers.reference_b(n)
    ks = [3, 30, 60]
    calls = []

    def ftk(al, be):
        calls.append((len(al), len(be)))
        return helpers.FTK["exp"](-al, be)  # any closure of the coefficients

    X = tpl.lanczos_two_pass_sweep(op, b, ks, ftk)
    assert calls == [(3, 2), (30, 29), (60, 59)]       # the closure sees the leading k coefficients, once per k
    for q, k in enumerate(ks):
        assert np.array_equal(X[q], tpl.lanczos_two_pass(op, b, k, ftk))
    Xd = tpl.lanczos_sweep(op, torch.from_numpy(b).cuda(), ks, "square")
    assert Xd.is_cuda
    for q, k in enumerate(ks):
        assert np.array_equal(Xd[q].cpu().numpy(), tpl.lanczos(op, b, k, "square"))
    # breakdown: diag(2, 3), b = e1 stops after one step (mod.rs:410-419); larger k reuse the one-step decomposition
    small = tpl.LinOp.from_dense(np.diag([2.0, 3.0]))
    Xb = tpl.lanczos_sweep(small, [1.0, 0.0], [1, 2], "inv")
    assert np.allclose(Xb, [[0.5, 0.0], [0.5, 0.0]], rtol=1e-15)
    with pytest.raises(tpl.LanczosError) as e:
        tpl.lanczos_sweep(small, [1.0, 0.0], [2, 0], "inv")
    assert e.value.kind == "Panic"
    with pytest.raises(tpl.LanczosError) as e:
        tpl.lanczos_two_pass_sweep(small, [0.0, 0.0], [2], "inv")
    assert e.value.kind == "InputError"
    with pytest.raises(tpl.LanczosError) as e:
        tpl.lanczos_sweep(op, b, ks, lambda al, be: np.ones(2))
    assert e.value.kind == "ParameterMismatch"
