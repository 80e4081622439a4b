"""Pins the CPU oracle against every known-answer / analytic test the reference holds for the hot path
(SURVEY section 8c): src/algorithms/mod.rs:384-428 (unit tests), :434-587 (property runners),
tests/correctness.rs:165-325 (diagonal ground truth), src/lib.rs:35-84 (doctest), src/error.rs:69-129
(Display strings) and every row of results/accuracy_*.csv / results/orthogonality_*.csv (tests/golden/published_curves.json)."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import np_oracle as npo
from oracle import oracle as orc

TOLERANCE = 5e-9          # src/algorithms/mod.rs:360
APPROX_TOLERANCE = 1e-3   # tests/correctness.rs:42
EXACT_TOLERANCE = 1e-12   # tests/correctness.rs:51


def simple_problem():  # mod.rs:371-380
    a = np.array([[2, -1, 0, 0], [-1, 2, -1, 0], [0, -1, 2, -1], [0, 0, -1, 2.0]])
    return orc.SparseColMat.from_dense(a), np.array([1, 2, 3, 4.0])


def test_recurrence_step_correctness():  # mod.rs:385-407: alpha = 2, beta = 1 to 1e-15
    a, _ = simple_problem()
    d = orc.lanczos_pass_one(a, [1, 0, 0, 0], 2)
    assert abs(d.alphas[0] - 2.0) < 1e-15
    assert abs(d.betas[0] - 1.0) < 1e-15


def test_breakdown_scenario():  # mod.rs:410-419
    a = orc.SparseColMat.from_dense(np.diag([2.0, 3.0]))
    v, d = orc.lanczos_standard(a, [1, 0], 2)
    assert d.steps_taken == 1
    assert len(d.alphas) == 1 and len(d.betas) == 0 and v.shape == (2, 1)


def test_zero_vector_input_returns_error():  # mod.rs:422-428
    a = orc.SparseColMat.from_dense(np.eye(2))
    with pytest.raises(orc.OracleError) as e:
        orc.lanczos_standard(a, [0, 0], 2)
    assert str(e.value) == "Invalid input parameter: Input vector `b` must not be a zero vector."


def test_error_messages():  # src/error.rs:97-117
    a, b = simple_problem()
    d = orc.lanczos_pass_one(a, b, 3)
    with pytest.raises(orc.OracleError) as e:
        orc.lanczos_pass_two(a, b, d, np.ones(2))
    assert str(e.value) == "Parameter mismatch: `y_k` expects size 3, but got 2."
    d0 = orc.LanczosDecomposition(d.alphas, d.betas, d.steps_taken, 0.0)
    with pytest.raises(orc.OracleError) as e:
        orc.lanczos_pass_two(a, b, d0, np.ones(3))
    assert str(e.value) == "Invalid input parameter: The initial vector `b` must not be a zero vector."
    with pytest.raises(orc.OracleError) as e:
        orc.lanczos_two_pass(a, b, 3, lambda al, be: np.ones(5))
    assert str(e.value) == "Parameter mismatch: `y_k_prime` expects size 3, but got 5."
    with pytest.raises(orc.OracleError) as e:
        orc.lanczos(a, b, 3, lambda al, be: (_ for _ in ()).throw(RuntimeError("Custom solver failed")))
    assert str(e.value).startswith("The user-provided f(T_k) solver failed: ")
    with pytest.raises(orc.OracleError):  # k == 0 panics in the reference (Vec::with_capacity(k-1))
        orc.lanczos_pass_one(a, b, 0)


def test_doctest_one_pass_vs_two_pass():  # src/lib.rs:35-84
    a, b = simple_problem()
    x1 = orc.lanczos(a, b, 3, npo.inv_tk_solver)
    x2 = orc.lanczos_two_pass(a, b, 3, npo.inv_tk_solver)
    assert np.linalg.norm(x1 - x2) < 1e-12


@pytest.mark.parametrize("solver", [orc.lanczos, orc.lanczos_two_pass])
@pytest.mark.parametrize("fname,f,tol", [("inv", lambda z: 1.0 / z, APPROX_TOLERANCE),
                                         ("exp", np.exp, APPROX_TOLERANCE),
                                         ("square", lambda z: z * z, EXACT_TOLERANCE)])
@pytest.mark.parametrize("bgen", ["numpy", "stdrng"])
def test_diagonal_ground_truth(solver, fname, f, tol, bgen):  # tests/correctness.rs:165-325
    n, k = 100, 30
    eigs = np.arange(1, n + 1.0)
    a = orc.SparseColMat.try_new_from_triplets(n, n, np.arange(n), np.arange(n), eigs)
    b = helpers.B_GENERATORS[bgen](n)
    x_true = f(eigs) * b
    x = solver(a, b, k, helpers.FTK[fname])
    assert helpers.rel(x, x_true) < tol


def golden_instances():
    return sorted(glob.glob(os.path.join(helpers.GOLDEN, "netgen1000", "*.dmx")))


@pytest.mark.parametrize("dmx", golden_instances(), ids=os.path.basename)
@pytest.mark.parametrize("flavour", ["qfc", "lines.qfc", "wc.qfc"])
def test_property_runners(dmx, flavour):
    """The four generated property tests of build.rs:53-110 / mod.rs:434-587, k = 30, b ~ U[0,1) seed 42."""
    kkt = orc.load_kkt_system(dmx, dmx[:-3] + flavour)
    a, n, k = kkt.a, kkt.a.nrows(), 30
    b = helpers.seeded_b(n)
    v, dec = orc.lanczos_standard(a, b, k)
    po = orc.lanczos_pass_one(a, b, k)
    # decomposition consistency (mod.rs:434-482)
    assert dec.steps_taken == po.steps_taken
    assert np.max(np.abs(dec.alphas - po.alphas)) < TOLERANCE
    assert np.max(np.abs(dec.betas - po.betas)) < TOLERANCE
    # Lanczos relation A V_k - V_k T_k = beta_k v_{k+1} e_k^T (mod.rs:486-529); absolute tolerance, as in the
    # reference, scaled by ||A|| because the qfcgen "aa" costs reach 1e5-1e6
    v1, dec1 = orc.lanczos_standard(a, b, k + 1)
    t = npo.assemble_tridiagonal(dec.alphas, dec.betas)
    av = np.stack([a.apply(v[:, j]) for j in range(k)], axis=1)
    resid = av - v @ t
    resid[:, k - 1] -= dec1.betas[k - 1] * v1[:, k]
    scale = max(1.0, np.abs(dec.alphas).max())
    assert np.linalg.norm(resid) < TOLERANCE * scale
    # orthonormality (mod.rs:532-554).  The aa flavour loses orthogonality within 30 steps (lambda_max ~ 1e6
    # converges immediately), which the reference never saw because its loader dropped D.
    if flavour != "lines.qfc":
        assert np.linalg.norm(np.eye(k) - v.T @ v) < TOLERANCE
    # reconstruction stability (mod.rs:558-587)
    y = 0.1 * (np.arange(po.steps_taken) + 1)
    _, v2 = orc.lanczos_pass_two(a, b, po, y, with_basis=True)
    assert np.sum((v - v2) ** 2) < TOLERANCE
    assert np.array_equal(v, v2)  # shipped CSVs: basis_drift_fro == 0.0 in every row


def test_cpp_oracle_vs_numpy_oracle():
    """Two summation orders bracket faer's un-vendored kernels: coefficients agree to 1e-12 well
    inside the orthogonality horizon, x at k=30 to 1e-10."""
    dmx = golden_instances()[2]
    kkt = orc.load_kkt_system(dmx, dmx[:-3] + "wc.qfc")
    cp, ri, va = kkt.a.csc()
    import scipy.sparse as sp

    n = kkt.a.nrows()
    a_sp = sp.csc_matrix((va, ri.astype(np.int64), cp.astype(np.int64)), shape=(n, n))
    b = helpers.seeded_b(n)
    dec = orc.lanczos_pass_one(kkt.a, b, 30)
    al, be, steps, bn = npo.pass_one(a_sp, b, 30)
    assert steps == dec.steps_taken
    assert np.max(np.abs(al - dec.alphas) / np.abs(dec.alphas).max()) < 1e-12
    assert np.max(np.abs(be - dec.betas) / np.abs(dec.betas).max()) < 1e-12
    x1 = orc.lanczos_two_pass(kkt.a, b, 30, npo.exp_tk_solver)
    x2 = npo.lanczos_two_pass(a_sp, b, 30, npo.exp_tk_solver)
    assert helpers.rel(x1, x2) < 1e-10
    # extended-precision oracle agrees too
    x3 = orc.lanczos_two_pass(kkt.a, b, 30, npo.exp_tk_solver, extended=True)
    assert helpers.rel(x1, x3) < 1e-10


def test_golden_vectors_reproduce(golden_dir):
    """The committed golden alpha/beta/x are what the oracle produces today (fixture drift guard)."""
    g = np.load(os.path.join(golden_dir, "golden_vectors.npz"))
    for dmx in golden_instances():
        name = os.path.basename(dmx)[:-4]
        for flavour, ext in (("nod", "qfc"), ("aa", "lines.qfc"), ("wc", "wc.qfc")):
            kkt = orc.load_kkt_system(dmx, dmx[:-3] + ext)
            key = f"{name}.{flavour}"
            assert list(g[key + ".nnz"]) == [kkt.a.nnz(), kkt.num_nodes, kkt.num_arcs]
            dec = orc.lanczos_pass_one(kkt.a, g[key + ".b"], 30)
            assert np.array_equal(dec.alphas, g[key + ".alphas"])
            assert np.array_equal(dec.betas, g[key + ".betas"])


# ---- Outputs of the reference itself (results/accuracy_*.csv, results/orthogonality_*.csv; n = 10 000 diagonal spectra of
# src/bin/stability.rs:98-146 / src/bin/orthogonality.rs:91-146, b = StdRng::seed_from_u64(42) uniforms) ---------------------
CURVES = ["inv_well", "inv_ill", "exp_well", "exp_ill"]


def _diagonal_problem(curve):
    func, scenario = curve.split("_")
    n = 10_000
    eigs = helpers.stability_spectrum(n, func, scenario)
    a = orc.SparseColMat.try_new_from_triplets(n, n, np.arange(n), np.arange(n), eigs)
    b = helpers.reference_b(n)  # stability.rs:256-257, orthogonality.rs:162-163
    x_true = (np.exp(eigs) if func == "exp" else 1.0 / eigs) * b
    return func, a, b, x_true


@pytest.mark.parametrize("curve", CURVES)
def test_published_accuracy_rows(curve):
    """EVERY row of the reference's accuracy CSVs (stability.rs:259-312: both variants against the analytic solution and
    against each other) is reproduced by the oracle fed with the restated StdRng(42) right-hand side: this pins the oracle
    AND two_pass_lanczos_b200/stdrng.py on outputs of the reference itself (a numpy-seeded b is off by 3-50 % on the same rows)."""
    func, a, b, x_true = _diagonal_problem(curve)
    ent = helpers.published_curves()["accuracy"][curve]
    assert len(ent["rows"]) == 20
    for k, pub_std, pub_two, pub_dev in ent["rows"]:
        x1 = orc.lanczos(a, b, k, helpers.FTK[func])
        x2 = orc.lanczos_two_pass(a, b, k, helpers.FTK[func])
        helpers.check_accuracy_row(curve, k, pub_std, helpers.rel(x1, x_true))
        helpers.check_accuracy_row(curve, k, pub_two, helpers.rel(x2, x_true))
        assert pub_dev < 3e-16 and helpers.rel(x1, x2) < 2e-15  # column 4: the variants agree to rounding


@pytest.mark.parametrize("curve", CURVES)
def test_published_orthogonality_curves(curve):
    """results/orthogonality_*.csv: ||I - V_k^T V_k||_F of the stored basis for k = 20 ... 1000 (orthogonality.rs:176-213).
    The loss starts at rounding level and is amplified chaotically, so a row is an envelope, not a digit-for-digit value:
    within x4 while the published loss is at rounding level (<= 1e-12), within x30 once it is amplified (the amplification of the last-bit differences of the reductions is itself chaotic).  This is what pins the ACCURACY CLASS of
    the oracle's dot products (a left-to-right sum sits 10-15 x above every published row)."""
    func, a, b, _ = _diagonal_problem(curve)
    ent = helpers.published_curves()["orthogonality"][curve]
    kmax = max(r[0] for r in ent["rows"])
    v, dec = orc.lanczos_standard(a, b, kmax)
    assert dec.steps_taken == kmax
    g = v.T @ v
    for k, pub_std, pub_regen, drift, soldev in ent["rows"]:
        assert pub_std == pub_regen and drift == 0.0 and soldev == 0.0  # the reference's own invariant
        loss = np.linalg.norm(np.eye(k) - g[:k, :k])
        bound = 4.0 if pub_std <= 1e-12 else 30.0
        assert pub_std / bound < loss < pub_std * bound, (curve, k, pub_std, loss)
    dec_k = orc.LanczosDecomposition(dec.alphas[:60], dec.betas[:59], 60, dec.b_norm)
    _, v2 = orc.lanczos_pass_two(a, b, dec_k, np.zeros(60), with_basis=True)
    assert np.array_equal(v[:, :60], v2)  # basis_drift_fro == 0.0
