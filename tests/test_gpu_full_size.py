"""BASELINE.json full-size configurations (500k arcs, rho = 3, k up to 500) checked through size-independent
properties, plus the accuracy / orthogonality harness of config 5 (diagonal spectra of src/bin/stability.rs and
src/bin/orthogonality.rs, n = 10 000)."""
import numpy as np
import pytest

import helpers
import two_pass_lanczos_b200 as tpl
from oracle import np_oracle as npo
from oracle import oracle as orc
from two_pass_lanczos_b200 import algorithms as alg
from two_pass_lanczos_b200 import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def headline():
    inst = datagen.gen_kkt(500_000, 3, 1, "aa")
    assert inst.n == 501_155  # results/scalability_k500_rho3.csv:20-21
    gop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    b = gop.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n)))  # src/bin/tradeoff.rs:234-236
    return inst, gop, b


def test_headline_two_pass_residual_and_variants(headline):
    inst, gop, b = headline
    k = 500
    x2 = tpl.lanczos_two_pass(gop, b, k, "inv")
    x1 = tpl.lanczos(gop, b, k, "inv")
    assert np.all(np.isfinite(x2))
    assert helpers.rel(x1, x2) < 1e-12  # accuracy_*.csv col 4
    res = np.linalg.norm(gop.apply(x2) - b) / np.linalg.norm(b)
    assert res < 1e-8
    # run-to-run determinism: fixed reduction order, no atomics
    assert np.array_equal(tpl.lanczos_two_pass(gop, b, k, "inv"), x2)
    # linearity in b: f(A)(2b) = 2 f(A) b  (power-of-two scaling commutes with every rounding)
    assert np.array_equal(tpl.lanczos_two_pass(gop, 2.0 * b, k, "inv"), 2.0 * x2)


def test_kernel_shapes_by_size(headline):
    """The headline instance runs on the cell kernels; 650k arcs (cells too large) on the chunk-resident or the tiled ones,
    whichever fits; 2M arcs on the blocked streaming ones (from ~1.5M arcs); every mode switch reports what it selects."""
    inst, gop, b = headline
    assert gop.kernel_shape() == "cells"
    for mode, shape in ((4, "chunks"), (5, "blocked"), (2, "tiled"), (3, "gather"), (1, "gather"), (0, "cells")):
        gop.set_mode(mode)
        assert gop.kernel_shape() == shape
    for m, shapes in ((650_000, ("chunks", "tiled")), (2_000_000, ("blocked",))):
        big = datagen.gen_kkt(m, 3, 2, "wc")
        op = tpl.LinOp.from_kkt(big.m, big.p, big.tail, big.head, big.d)
        assert op.kernel_shape() in shapes, op.kernel_shape()
        bb = op.apply(np.full(big.n, 1.0 / np.sqrt(big.n)))
        d1 = alg.lanczos_pass_one(op, bb, 30)
        op.set_mode(3)  # the gather kernels as an independent implementation
        d2 = alg.lanczos_pass_one(op, bb, 30)
        assert np.max(np.abs(d1.alphas - d2.alphas)) <= 1e-10 * np.abs(d2.alphas).max()
        assert np.max(np.abs(d1.betas - d2.betas)) <= 1e-10 * np.abs(d2.betas).max()
        op.close()


def test_headline_against_oracle_prefix(headline):
    """alpha/beta vs the CPU oracle on the full-size instance for the first 40 steps (seconds on one core)."""
    inst, gop, b = headline
    oop = helpers.oracle_op(inst)
    k = 40
    d_ref = orc.lanczos_pass_one(oop, b, k)
    d = alg.lanczos_pass_one(gop, b, k)
    assert d.steps_taken == d_ref.steps_taken == k
    J = 12  # lambda_max ~ 1e6 converges almost at once on the aa flavour; orthogonality is lost soon after
    assert np.max(np.abs(d.alphas[:J] - d_ref.alphas[:J])) <= 1e-12 * np.abs(d_ref.alphas).max()
    assert np.max(np.abs(d.betas[:J] - d_ref.betas[:J])) <= 1e-12 * np.abs(d_ref.betas).max()
    y = np.zeros(k)
    y[:J] = 1.0 / (1.0 + np.arange(J))
    assert helpers.rel(alg.lanczos_pass_two(gop, b, d, y), orc.lanczos_pass_two(oop, b, d_ref, y)) < 1e-9


def test_headline_basis_regeneration_is_exact(headline):
    inst, gop, b = headline
    k = 24
    out = alg.lanczos_standard(gop, b, k)
    p2 = alg.lanczos_pass_two_with_basis(gop, b, out.decomposition, np.zeros(k))
    assert np.array_equal(out.v_k, p2.v_k)


def test_wc_full_size_converged_solution_matches_oracle():
    """500k arcs, wc flavour, f = inv, k = 250 (inside the plateau): projected x within 1e-10 of the
    numpy oracle (different summation order) and of the known solution."""
    inst = datagen.gen_kkt(500_000, 3, 2, "wc")
    gop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    a_sp = npo.kkt_matrix(inst.m, inst.p, inst.tail.astype(np.int64), inst.head.astype(np.int64), inst.d)
    x_true = np.full(inst.n, 1.0 / np.sqrt(inst.n))
    b = a_sp @ x_true
    k = 250
    x = helpers.project_out_null(tpl.lanczos_two_pass(gop, b, k, "inv"), inst.m, inst.p)
    x_ref = helpers.project_out_null(npo.lanczos_two_pass(a_sp, b, k, npo.inv_tk_solver), inst.m, inst.p)
    assert helpers.rel(x, x_ref) < 1e-10
    assert helpers.rel(x, helpers.project_out_null(x_true, inst.m, inst.p)) < 1e-10


def test_headline_k500_matches_oracle(headline):
    """THE north-star acceptance gate: `lanczos_two_pass`, f = inv, 500 k arcs "aa", rho = 3, k = 500, b = A (1/sqrt n) 1
    (src/bin/tradeoff.rs:234-258, src/solvers.rs:133-175): x within 1e-10 relative of the CPU reference path (one oracle
    solve, a few seconds on one core).  SURVEY 8c(ii): on the aa flavour the null direction is not found by k = 500, so the
    raw x is compared; the null-space-projected deviation and the residual are checked beside it."""
    inst, gop, b = headline
    oop = helpers.oracle_op(inst)
    b_cpu = helpers.rhs_from_const(oop.apply, inst.n)
    assert helpers.rel(b, b_cpu) < 1e-15
    k = 500
    x_cpu = orc.lanczos_two_pass(oop, b_cpu, k, npo.inv_tk_solver)
    x_gpu = tpl.lanczos_two_pass(gop, b, k, "inv")
    assert helpers.rel(x_gpu, x_cpu) <= 1e-10
    assert helpers.rel(helpers.project_out_null(x_gpu, inst.m, inst.p), helpers.project_out_null(x_cpu, inst.m, inst.p)) <= 1e-10
    x1_gpu = tpl.lanczos(gop, b, k, "inv")
    assert helpers.rel(x1_gpu, x_cpu) <= 1e-10
    for shape_mode in (5, 2, 3):  # the blocked and the tiled streaming kernels and the gather kernels on the same instance
        gop.set_mode(shape_mode)
        assert helpers.rel(tpl.lanczos_two_pass(gop, b, k, "inv"), x_cpu) <= 1e-10, shape_mode
    gop.set_mode(0)


@pytest.mark.parametrize("fmt", ["kkt", "csr"])
def test_config2_50k_k500_one_pass_vs_two_pass_matches_oracle(fmt):
    """BASELINE config 2: 50 k arcs, rho = 3, k = 500, one-pass vs two-pass (results/tradeoff_arcs50k_rho3.csv:11,31).
    (a) "wc" costs (the well-conditioned flavour, tex/report.tex:338-342) inside the convergence plateau: x within 1e-10 of the
        oracle, raw and null-space projected, for both variants.
    (b) the qfcgen "aa" costs at k = 500, the published point: kappa ~ 1e8 and the iteration is NOT converged there (error vs
        the known solution 4.7e-4), so x depends on the last bit of the summation order -- two CPU evaluations of the same
        reference path differ by 2e-5.  The GPU result must lie inside that spread (x10), reach the same error against the
        known solution, and its two variants must agree to rounding."""
    def make(inst, oop):
        if fmt == "kkt":
            return tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
        return tpl.LinOp.from_csc(inst.n, *oop.csc())

    proj = helpers.project_out_null
    # (a)
    inst = datagen.gen_kkt(50_000, 3, 1, "wc")
    assert inst.n == 50_365
    oop = helpers.oracle_op(inst)
    gop = make(inst, oop)
    b = helpers.rhs_from_const(oop.apply, inst.n)
    x_cpu = orc.lanczos_two_pass(oop, b, 200, npo.inv_tk_solver)
    x2, x1 = tpl.lanczos_two_pass(gop, b, 200, "inv"), tpl.lanczos(gop, b, 200, "inv")
    for x in (x1, x2):
        assert helpers.rel(x, x_cpu) <= 1e-10
        assert helpers.rel(proj(x, inst.m, inst.p), proj(x_cpu, inst.m, inst.p)) <= 1e-10
    gop.close()
    # (b)
    inst = datagen.gen_kkt(50_000, 3, 1, "aa")
    oop = helpers.oracle_op(inst)
    gop = make(inst, oop)
    b = helpers.rhs_from_const(oop.apply, inst.n)
    x_cpu = orc.lanczos_two_pass(oop, b, 500, npo.inv_tk_solver)
    spread = helpers.cpu_spread(inst, b, 500, x_cpu)
    x2, x1 = tpl.lanczos_two_pass(gop, b, 500, "inv"), tpl.lanczos(gop, b, 500, "inv")
    assert helpers.rel(x2, x_cpu) <= max(1e-10, 10.0 * spread), (helpers.rel(x2, x_cpu), spread)
    assert helpers.rel(x1, x2) <= 1e-12
    x_true = proj(np.full(inst.n, 1.0 / np.sqrt(inst.n)), inst.m, inst.p)
    e_gpu, e_cpu = helpers.rel(proj(x2, inst.m, inst.p), x_true), helpers.rel(proj(x_cpu, inst.m, inst.p), x_true)
    assert 0.5 * e_cpu <= e_gpu <= 2.0 * e_cpu, (e_gpu, e_cpu)
    gop.close()


@pytest.mark.parametrize("k", [50, 250, 1000])
def test_config3_k_sweep_points_match_oracle(headline, k):
    """BASELINE config 3 (k = 50 ... 1000 at 500 k arcs, results/tradeoff_arcs500k_rho3.csv): x vs the oracle at three more
    points of the sweep (k = 500 is the headline gate above).  Unconverged points are gated by the spread two CPU
    evaluations of the reference path show on the same problem (helpers.cpu_spread), converged ones by 1e-10."""
    inst, gop, b = headline
    oop = helpers.oracle_op(inst)
    x_cpu = orc.lanczos_two_pass(oop, b, k, npo.inv_tk_solver)
    x_gpu = tpl.lanczos_two_pass(gop, b, k, "inv")
    spread = helpers.cpu_spread(inst, b, k, x_cpu)
    assert helpers.rel(x_gpu, x_cpu) <= max(1e-10, 10.0 * spread), (k, helpers.rel(x_gpu, x_cpu), spread)


CURVES = ["inv_well", "inv_ill", "exp_well", "exp_ill"]


@pytest.mark.parametrize("curve", CURVES)
def test_stability_harness(curve, tmp_path):
    """src/bin/stability.rs on the GPU path (two_pass_lanczos_b200/experiments.py, the code behind scripts/stability.py): EVERY
    row of the reference's results/accuracy_*.csv (both variants against the analytic f(lambda_i) b_i, b =
    StdRng::seed_from_u64(42) uniforms) within the per-curve tolerance of tests/helpers.py, the two variants equal to rounding
    (column 4), and the CSV written with the reference's schema."""
    from two_pass_lanczos_b200 import experiments

    func, scenario = curve.split("_")
    pub = helpers.published_curves()["accuracy"][curve]["rows"]
    assert len(pub) == 20
    rows = experiments.run_accuracy(func, scenario + "-conditioned", n=10_000, k_min=10, k_max=200, k_step=10)
    assert [r[0] for r in rows] == [r[0] for r in pub]
    for (k, e1, e2, dev), (_, pub_std, pub_two, _) in zip(rows, pub):
        helpers.check_accuracy_row(curve, k, pub_std, e1)
        helpers.check_accuracy_row(curve, k, pub_two, e2)
        assert dev < 2e-15
    out = tmp_path / "acc.csv"
    experiments.write_csv(str(out), experiments.ACCURACY_COLUMNS, rows)
    lines = out.read_text().splitlines()
    assert lines[0] == "k,relative_error_standard,relative_error_two_pass,relative_solution_deviation" and len(lines) == 21


@pytest.mark.parametrize("curve", CURVES)
def test_orthogonality_harness(curve):
    """src/bin/orthogonality.rs:148-232 on the GPU path (experiments.run_orthogonality, behind scripts/orthogonality.py):
    ||I - V^T V||_F of the stored and of the regenerated basis, and their drift, against the reference's
    results/orthogonality_*.csv (drift exactly 0.0, both losses bit-identical; the loss itself is an envelope: x4 at rounding
    level, x30 once it is amplified (the amplification of the last-bit differences of the reductions is itself chaotic) -- same bounds as for the oracle)."""
    from two_pass_lanczos_b200 import experiments

    func, scenario = curve.split("_")
    pub = {r[0]: r for r in helpers.published_curves()["orthogonality"][curve]["rows"]}
    rows = experiments.run_orthogonality(func, scenario + "-conditioned", n=10_000, k_min=100, k_max=1000, k_step=300)
    rows += experiments.run_orthogonality(func, scenario + "-conditioned", n=10_000, k_min=20, k_max=20, k_step=20)
    assert sorted(r[0] for r in rows) == [20, 100, 400, 700, 1000]
    for k, loss_std, loss_regen, drift, soldev in rows:
        assert drift == 0.0 and soldev == 0.0      # basis_drift_fro, solution_deviation_l2: exactly zero, as published
        assert loss_std == loss_regen
        bound = 4.0 if pub[k][1] <= 1e-12 else 30.0
        assert pub[k][1] / bound < loss_std < pub[k][1] * bound, (curve, k, pub[k][1], loss_std)


def test_csr_path_at_5m_arcs_matches_oracle_and_incidence_path():
    """north_star "Matvec": the generic CSR kernels (SELL-32 slices + long-row segments, tpl_csr.cuh) on the matrix exactly as
    `load_kkt_system` hands it back (KKTSystem.a, src/utils/data_loader.rs:251) at 5M arcs: alpha / beta against the CPU oracle
    up to the orthogonality horizon of the wc flavour, x against the incidence path of the same operator, residual."""
    inst = datagen.gen_kkt(5_000_000, 3, 1, "wc")
    cp, ri, va = datagen.kkt_csc(inst)
    gop = tpl.LinOp.from_csc(inst.n, cp, ri, va)
    assert gop.format == "csr" and gop.matrix_bytes() == 12 * len(va) + 4 * (inst.n + 1)
    oop = helpers.oracle_op(inst)
    b = helpers.rhs_from_const(oop.apply, inst.n)
    assert helpers.rel(gop.apply(np.full(inst.n, 1.0 / np.sqrt(inst.n))), b) < 1e-14
    J = 16
    d_ref = orc.lanczos_pass_one(oop, b, J)
    d = alg.lanczos_pass_one(gop, b, J)
    assert d.steps_taken == d_ref.steps_taken == J
    assert np.max(np.abs(d.alphas - d_ref.alphas)) <= 1e-12 * np.abs(d_ref.alphas).max()
    assert np.max(np.abs(d.betas - d_ref.betas)) <= 1e-12 * np.abs(d_ref.betas).max()
    iop = tpl.LinOp.from_kkt(inst.m, inst.p, inst.tail, inst.head, inst.d)
    assert iop.kernel_shape() == "blocked"
    k = 120
    x_csr = tpl.lanczos_two_pass(gop, b, k, "inv")
    x_inc = tpl.lanczos_two_pass(iop, b, k, "inv")
    proj = helpers.project_out_null
    assert helpers.rel(proj(x_csr, inst.m, inst.p), proj(x_inc, inst.m, inst.p)) <= 1e-10
    assert np.linalg.norm(gop.apply(x_csr) - b) / np.linalg.norm(b) < 1e-7
    gop.close()
    iop.close()
