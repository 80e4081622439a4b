"""Host-side loader (tpl_load_kkt, C++) against the oracle's restatement of src/utils/data_loader.rs, on the
reference tools' own output (tests/golden/netgen1000) and on every DataLoaderError branch.  No GPU needed:
the loader is host code behind the C ABI."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import oracle as orc
from two_pass_lanczos_b200 import data_loader, datagen
from two_pass_lanczos_b200.error import DataLoaderError


def golden_instances():
    return sorted(glob.glob(os.path.join(helpers.GOLDEN, "netgen1000", "*.dmx")))


@pytest.mark.parametrize("dmx", golden_instances(), ids=os.path.basename)
@pytest.mark.parametrize("ext,has_d", [("qfc", False), ("lines.qfc", True), ("wc.qfc", True)])
def test_loader_matches_oracle_on_reference_tool_output(dmx, ext, has_d):
    qfc = dmx[:-3] + ext
    host = data_loader.load_kkt_host(dmx, qfc)
    ref = orc.load_kkt_system(dmx, qfc)
    n, colptr, rowidx, val = host.csc()
    rc, rr, rv = ref.a.csc()
    assert (host.num_nodes, host.num_arcs) == (ref.num_nodes, ref.num_arcs)
    assert n == ref.a.nrows()
    assert np.array_equal(colptr, rc) and np.array_equal(rowidx, rr) and np.array_equal(val, rv)
    m = host.num_arcs
    # qfcgen's 3-line file makes the reference loader drop D -> nnz = 4m; one value per line -> 5m
    assert host.num_costs == (m if has_d else 0)
    assert len(val) == (5 * m if has_d else 4 * m)
    tail, head, d, d_len, regular = host.incidence()
    assert regular and d_len == host.num_costs
    # incidence view reproduces the CSC exactly
    import scipy.sparse as sp

    a = sp.csc_matrix((val, rowidx.astype(np.int64), colptr.astype(np.int64)), shape=(n, n)).toarray()
    j = np.arange(m)
    b = np.zeros((n, n))
    b[j[:d_len], j[:d_len]] = d[:d_len]
    b[m + tail, j] += 1.0
    b[m + head, j] -= 1.0
    b[j, m + tail] += 1.0
    b[j, m + head] -= 1.0
    assert np.array_equal(a, b)
    assert np.array_equal(a, a.T)


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text if isinstance(text, bytes) else text.encode())
    return str(p)


GOOD_DMX = "c comment\np min 3 2\nn 1 5\na 1 2 0 9 9\na 2 3 0 9 9\n"
GOOD_QFC = "2\n1.0\n1.0\n3.5\n4.5\n"

CASES = [  # (dmx text, qfc text, error kind, message)
    (None, GOOD_QFC, "Io", None),
    (GOOD_DMX, None, "Io", None),
    ("c only comments\na 1 2\n", GOOD_QFC, "ProblemLineMissing",
     "Format error: The 'p min' problem line was not found or was malformed."),
    ("p max 3 2\n", GOOD_QFC, "ProblemLineMissing",
     "Format error: The 'p min' problem line was not found or was malformed."),
    ("p min 3\n", GOOD_QFC, "ProblemLineMissing",
     "Format error: The 'p min' problem line was not found or was malformed."),
    ("p min x3 2\n", GOOD_QFC, "ParseInt", "Parse error: Failed to parse integer from 'x3'"),
    ("p min 3 -2\n", GOOD_QFC, "ParseInt", "Parse error: Failed to parse integer from '-2'"),
    ("p min 3 2\na 1 two\n", GOOD_QFC, "ParseInt", "Parse error: Failed to parse integer from 'two'"),
    ("p min 3 2\na 0 2\n", GOOD_QFC, "InvalidDimacsNodeIndex",
     "Format error: Invalid node index '0'. DIMACS format requires 1-based positive integers."),
    ("p min 3 2\na 1 4\na 2 3\n", GOOD_QFC, "SparseMatrixConstructionError",
     "Internal error: Failed to construct the sparse matrix from triplets."),
    ("p min 3 1\na 1 2\na 2 3\n", "1\n1\n2\n", "SparseMatrixConstructionError",
     "Internal error: Failed to construct the sparse matrix from triplets."),
    (GOOD_DMX, "", "UnexpectedEof", "Format error: Unexpected end of file while reading data."),
    (GOOD_DMX, "2 \n1\n1\n1\n1\n", "ParseInt", "Parse error: Failed to parse integer from 'm'"),
    (GOOD_DMX, "3\n1\n1\n1\n1\n", "ArcCountMismatch",
     "Dimension mismatch: qfc file specifies 3 arcs, but dmx file has 2."),
    (GOOD_DMX, "2\n1\n1\n3.5\nabc\n", "ParseFloat", "Parse error: Failed to parse float from 'abc'"),
    (GOOD_DMX, "2\n1\n1\n3.5 \n4\n", "ParseFloat", "Parse error: Failed to parse float from '3.5 '"),
    (GOOD_DMX, "2\n1\n1\n0x10\n4\n", "ParseFloat", "Parse error: Failed to parse float from '0x10'"),
    (GOOD_DMX.encode() + b"c \xff\xfe\n", GOOD_QFC, "Io", None),
    # BufRead::lines strips '\r' only as part of "\r\n": a last line WITHOUT '\n' keeps it and then fails to parse
    (GOOD_DMX, "2\r", "ParseInt", "Parse error: Failed to parse integer from 'm'"),
    (GOOD_DMX, "2\n1\n1\n3.5\n4.5\r", "ParseFloat", "Parse error: Failed to parse float from '4.5\r'"),
    # core::str::from_utf8 rejects overlong forms, surrogates and code points above U+10FFFF
    (GOOD_DMX.encode() + b"c \xc0\x80\n", GOOD_QFC, "Io", None),
    (GOOD_DMX.encode() + b"c \xe0\x80\x80\n", GOOD_QFC, "Io", None),
    (GOOD_DMX.encode() + b"c \xed\xa0\x80\n", GOOD_QFC, "Io", None),
    (GOOD_DMX.encode() + b"c \xf0\x80\x80\x80\n", GOOD_QFC, "Io", None),
    (GOOD_DMX.encode() + b"c \xf4\x90\x80\x80\n", GOOD_QFC, "Io", None),
    (GOOD_DMX.encode() + b"c \xf5\x80\x80\x80\n", GOOD_QFC, "Io", None),
]


def test_loader_accepts_valid_multibyte_utf8_and_crlf_last_line(tmp_path):
    """The strict UTF-8 check still accepts every well-formed sequence (2-, 3-, 4-byte, U+D7FF, U+E000, U+10FFFF), and a
    "\r\n"-terminated last line loses its '\r' as in BufRead::lines."""
    ok = "c \u00e9 \u20ac \ud7ff \ue000 \U0001f600 \U0010ffff\n".encode()
    dmx = _write(tmp_path, "u.dmx", ok + GOOD_DMX.encode())
    qfc = _write(tmp_path, "u.qfc", "2\r\n1.0\r\n1.0\r\n3.5\r\n4.5\r\n")
    host = data_loader.load_kkt_host(dmx, qfc)
    ref = orc.load_kkt_system(dmx, qfc)
    assert host.num_costs == 2 and np.array_equal(host.csc()[3], ref.a.csc()[2])


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_loader_error_branches(tmp_path, idx):
    dmx_t, qfc_t, kind, msg = CASES[idx]
    dmx = _write(tmp_path, "x.dmx", dmx_t) if dmx_t is not None else str(tmp_path / "missing.dmx")
    qfc = _write(tmp_path, "x.qfc", qfc_t) if qfc_t is not None else str(tmp_path / "missing.qfc")
    with pytest.raises(DataLoaderError) as e:
        data_loader.load_kkt_host(dmx, qfc)
    assert e.value.kind == kind
    if msg:
        assert str(e.value) == msg
    with pytest.raises(orc.OracleError) as eo:  # the oracle takes the same branch with the same text
        orc.load_kkt_system(dmx, qfc)
    assert eo.value.code == e.value.code
    if msg:
        assert str(eo.value) == msg


def test_loader_quirks(tmp_path):
    """Accepted oddities of the reference parser: CRLF, '+' signs, extra tokens, ignored line kinds, a later
    `p` line overriding an earlier one, short D, self-loops (merged explicit zero), duplicate arcs, inf/nan."""
    dmx = _write(tmp_path, "q.dmx",
                 "c x\r\np min 9 9\r\np min 4 5 extra\r\nn 1 3\r\nx y z\r\n\r\n"
                 "a +1 2 junk\r\na 2 2\r\na 3 4\r\na 3 4\r\na 4 1\r\n")
    qfc = _write(tmp_path, "q.qfc", "5\r\n-\r\nskipped\r\n\r\n?\r\n5\r\n+1.5\r\n.5\r\n2.\r\n1e3\r\ninf\r\nIGNORED\r\n")
    host = data_loader.load_kkt_host(dmx, qfc)
    ref = orc.load_kkt_system(dmx, qfc)
    n, colptr, rowidx, val = host.csc()
    rc, rr, rv = ref.a.csc()
    assert (host.num_nodes, host.num_arcs, n) == (4, 5, 9)
    assert np.array_equal(colptr, rc) and np.array_equal(rowidx, rr) and np.array_equal(val, rv)
    tail, head, d, d_len, regular = host.incidence()
    assert regular and d_len == 5
    assert list(d) == [1.5, 0.5, 2.0, 1000.0, np.inf]
    assert list(tail) == [0, 1, 2, 2, 3] and list(head) == [1, 1, 3, 3, 0]
    # the self-loop (arc 1) is one explicit 0.0 in E, so A holds explicit zeros at (m+1, 1) and (1, m+1)
    assert 0.0 in val
    # short D: fewer than m quadratic lines is accepted silently (data_loader.rs:187-195)
    qfc2 = _write(tmp_path, "s.qfc", "5\n1\n1\n1\n1\n1\n7.0\n8.0\n")
    host2 = data_loader.load_kkt_host(dmx, qfc2)
    assert host2.num_costs == 2
    assert len(host2.csc()[3]) == len(val) - 3
    # fewer `a` lines than the p line announces is only a debug_assert in the reference (data_loader.rs:145-148)
    dmx3 = _write(tmp_path, "f.dmx", "p min 3 4\na 1 2\na 2 3\n")
    host3 = data_loader.load_kkt_host(dmx3, _write(tmp_path, "f.qfc", "4\n0\n0\n0\n0\n1\n2\n3\n4\n"))
    assert not host3.incidence()[4]  # not a plain arc list -> generic CSR operator
    ref3 = orc.load_kkt_system(dmx3, str(tmp_path / "f.qfc"))
    assert np.array_equal(host3.csc()[3], ref3.a.csc()[2])


def test_malformed_arc_line_is_an_error_not_a_crash(tmp_path):
    """The reference panics (index out of bounds, data_loader.rs:118-119); the C ABI cannot unwind, so it
    reports a dedicated status instead."""
    dmx = _write(tmp_path, "m.dmx", "p min 3 2\na 1\n")
    with pytest.raises(DataLoaderError) as e:
        data_loader.load_kkt_host(dmx, _write(tmp_path, "m.qfc", GOOD_QFC))
    assert e.value.kind == "MalformedArcLine"


def test_generator_roundtrip_through_loader(tmp_path):
    """Seeded generator -> files -> loader returns exactly the generated arrays (both .qfc layouts)."""
    inst = datagen.gen_kkt(5000, 3, 3, "aa")
    assert inst.p == 115 and inst.n == 5115  # results/tradeoff_arcs5k_rho3.csv instance size
    dmx, qfc = datagen.write_instance(str(tmp_path), inst)
    assert os.path.basename(dmx) == "netgen-5000-3-3-a-a-ns.dmx"  # src/bin/datagen.rs:109-117
    host = data_loader.load_kkt_host(dmx, qfc)
    tail, head, d, d_len, regular = host.incidence()
    assert regular and d_len == inst.m
    assert np.array_equal(tail, inst.tail) and np.array_equal(head, inst.head) and np.array_equal(d, inst.d)
    cp, ri, va = datagen.kkt_csc(inst)
    n, colptr, rowidx, val = host.csc()
    assert np.array_equal(cp, colptr) and np.array_equal(ri, rowidx) and np.array_equal(va, val)
    q3 = str(tmp_path / "three.qfc")
    datagen.write_qfc(q3, inst, layout="qfcgen")
    assert data_loader.load_kkt_host(dmx, q3).num_costs == 0


def test_generator_shape():
    """netgen-shaped: p from pargen's formula, arcs grouped by increasing tail, sinks without out-arcs, sources
    without in-arcs, no self-loops, no duplicate (tail, head)."""
    for m, p in ((1000, 52), (50_000, 365), (500_000, 1155)):
        assert datagen.num_nodes(m, 3) == p
    assert datagen.num_nodes(1000, 1) == 89 and datagen.num_nodes(1000, 2) == 63
    a = datagen.gen_kkt(50_000, 3, 11, "aa")
    b = datagen.gen_kkt(50_000, 3, 11, "aa")
    assert np.array_equal(a.tail, b.tail) and np.array_equal(a.head, b.head) and np.array_equal(a.d, b.d)
    assert np.all(np.diff(a.tail.astype(np.int64)) >= 0)
    assert not np.any(a.tail == a.head)
    pairs = a.tail.astype(np.int64) * a.p + a.head
    assert len(np.unique(pairs)) == a.m
    assert a.d.min() >= 400.0 and a.d.max() <= 1.1e6
    w = datagen.gen_kkt(50_000, 3, 11, "wc")
    assert 1.0 <= w.d.min() and w.d.max() <= 10.0


# ------------------------------------------------------------------ binary container
def _same_system(a, b):
    assert (a.num_nodes, a.num_arcs, a.num_costs) == (b.num_nodes, b.num_arcs, b.num_costs)
    for x, y in zip(a.incidence(), b.incidence()):
        assert np.array_equal(x, y)
    na, ca, ra, va = a.csc()
    nb, cb, rb, vb = b.csc()
    assert na == nb and np.array_equal(ca, cb) and np.array_equal(ra, rb)
    assert np.array_equal(va, vb) and np.array_equal(np.signbit(va), np.signbit(vb))


@pytest.mark.parametrize("dmx", golden_instances()[:2], ids=os.path.basename)
@pytest.mark.parametrize("ext", ["qfc", "lines.qfc"])
def test_binary_container_round_trip_on_reference_tool_output(tmp_path, dmx, ext):
    """text pair -> container -> KKTSystem: incidence view and the CSC of KKTSystem.a (built straight from the arc
    list, no triplet sort) are entry for entry what the text path gives."""
    host = data_loader.load_kkt_host(dmx, dmx[:-3] + ext)
    path = str(tmp_path / "inst.tplkkt")
    host.save_binary(path)
    m, nd = host.num_arcs, host.num_costs
    assert os.path.getsize(path) == 64 + 2 * ((4 * m + 7) // 8 * 8) + 8 * nd
    _same_system(data_loader.load_kkt_host_binary(path), host)


def test_binary_container_quirks_survive(tmp_path):
    """self-loop (merged explicit zero), parallel arcs, short D, inf: the container keeps all of them"""
    dmx = _write(tmp_path, "q.dmx", "p min 4 5\na 1 2\na 2 2\na 3 4\na 3 4\na 4 1\n")
    qfc = _write(tmp_path, "q.qfc", "5\n1\n1\n1\n1\n1\n1.5\n-0.0\ninf\n")
    host = data_loader.load_kkt_host(dmx, qfc)
    assert host.num_costs == 3
    path = str(tmp_path / "q.tplkkt")
    host.save_binary(path)
    back = data_loader.load_kkt_host_binary(path)
    _same_system(back, host)
    # straight from arrays (what a generator would call), odd arc count -> padded index blocks
    tail, head, d, d_len, _ = host.incidence()
    path2 = str(tmp_path / "q2.tplkkt")
    data_loader.write_kkt_binary(path2, 4, tail, head, d[:d_len])
    assert open(path, "rb").read() == open(path2, "rb").read()
    # empty instance
    path3 = str(tmp_path / "e.tplkkt")
    data_loader.write_kkt_binary(path3, 0, [], [])
    e = data_loader.load_kkt_host_binary(path3)
    assert (e.num_nodes, e.num_arcs, e.csc()[0]) == (0, 0, 0)
    # an instance whose incidence view is not exact cannot be stored
    dmx3 = _write(tmp_path, "f.dmx", "p min 3 4\na 1 2\na 2 3\n")
    host3 = data_loader.load_kkt_host(dmx3, _write(tmp_path, "f.qfc", "4\n0\n0\n0\n0\n1\n2\n3\n4\n"))
    with pytest.raises(DataLoaderError) as err:
        host3.save_binary(str(tmp_path / "f.tplkkt"))
    assert err.value.kind == "SparseMatrixConstructionError"


def test_binary_container_generated_instance_matches_text_path(tmp_path):
    inst = datagen.gen_kkt(20_000, 3, 5, "wc")
    dmx, qfc = datagen.write_instance(str(tmp_path), inst, layout="lines")
    host = data_loader.load_kkt_host(dmx, qfc)
    path = str(tmp_path / "g.tplkkt")
    data_loader.write_kkt_binary(path, inst.p, inst.tail, inst.head, inst.d)
    _same_system(data_loader.load_kkt_host_binary(path), host)


def test_binary_container_error_branches(tmp_path):
    good = str(tmp_path / "g.tplkkt")
    data_loader.write_kkt_binary(good, 3, [0, 1, 2], [1, 2, 0], [1.0, 2.0])
    raw = open(good, "rb").read()

    def load(data):
        p = str(tmp_path / "bad.tplkkt")
        open(p, "wb").write(data)
        with pytest.raises(DataLoaderError) as e:
            data_loader.load_kkt_host_binary(p)
        return e.value.kind

    assert load(b"") == "UnexpectedEof"
    assert load(b"X" + raw[1:]) == "ProblemLineMissing"       # magic
    assert load(raw[:-1]) == "UnexpectedEof"                   # truncated payload
    assert load(raw + b"\0") == "ArcCountMismatch"             # trailing bytes
    flipped = bytearray(raw)
    flipped[64] ^= 1                                           # tail[0]: 0 -> 1, checksum no longer matches
    assert load(bytes(flipped)) == "Io"
    with pytest.raises(DataLoaderError) as e:
        data_loader.load_kkt_host_binary(str(tmp_path / "missing.tplkkt"))
    assert e.value.kind == "Io"
    # a header announcing 2^31 - 1 arcs on a 100-byte file is refused before anything is allocated
    import struct

    huge = raw[:16] + struct.pack("<QQQ", 3, 2**31 - 1, 0) + raw[40:]
    assert load(huge) == "UnexpectedEof"
    assert load(raw[:16] + struct.pack("<QQQ", 3, 2**31, 0) + raw[40:]) == "ArcCountMismatch"  # beyond the 32-bit index range
    with pytest.raises(DataLoaderError) as e:  # node id out of range is refused at write time, like the triplet check
        data_loader.write_kkt_binary(str(tmp_path / "o.tplkkt"), 3, [0, 3], [1, 2])
    assert e.value.kind == "SparseMatrixConstructionError"
    with pytest.raises(DataLoaderError) as e:
        data_loader.write_kkt_binary(str(tmp_path / "o.tplkkt"), 3, [0], [1], [1.0, 2.0])
    assert e.value.kind == "ArcCountMismatch"


def test_cost_spellings_round_like_the_reference(tmp_path):
    """<f64 as FromStr> corner spellings: saturating overflow / underflow, subnormals, long mantissas, exponent forms"""
    vals = ["1e400", "-1e400", "1e-400", "4.9e-324", "2.2250738585072011e-308", "0.1", "123456789012345678901234567890",
            "1E5", "-.5e-3", "+7.", "9007199254740993", "0.3000000000000000444089209850062616169452667236328125", "NaN", "-inf"]
    m = len(vals)
    dmx = _write(tmp_path, "v.dmx", "p min 2 %d\n" % m + "a 1 2\n" * m)
    qfc = _write(tmp_path, "v.qfc", "%d\n" % m + "0\n" * m + "\n".join(vals) + "\n")
    host = data_loader.load_kkt_host(dmx, qfc)
    ref = orc.load_kkt_system(dmx, qfc)
    d = host.incidence()[2]
    want = np.array([float(v) for v in vals])
    assert np.array_equal(d, want, equal_nan=True) and np.array_equal(np.signbit(d), np.signbit(want))
    assert np.array_equal(host.csc()[3], ref.a.csc()[2], equal_nan=True)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_direct_csc_construction_matches_the_oracles_triplet_route(tmp_path, seed):
    """random multigraph with self-loops, parallel arcs, isolated nodes and a short D: the CSC the product builds straight
    from the arc list equals the oracle's `try_new_from_triplets` restatement entry for entry"""
    rng = np.random.default_rng(seed)
    p, m = 60, 2500
    tail = rng.integers(1, p + 1, m)
    head = rng.integers(1, p + 1, m)
    head[::97] = tail[::97]  # self-loops
    n_costs = m - 7 * seed
    dmx = _write(tmp_path, "r.dmx", f"p min {p} {m}\n" + "".join(f"a {u} {v}\n" for u, v in zip(tail, head)))
    costs = rng.uniform(-5, 5, n_costs)
    qfc = _write(tmp_path, "r.qfc", f"{m}\n" + "0\n" * m + "".join(repr(float(c)) + "\n" for c in costs))
    host = data_loader.load_kkt_host(dmx, qfc)
    ref = orc.load_kkt_system(dmx, qfc)
    n, colptr, rowidx, val = host.csc()
    rc, rr, rv = ref.a.csc()
    assert n == m + p and host.num_costs == n_costs
    assert np.array_equal(colptr, rc) and np.array_equal(rowidx, rr) and np.array_equal(val, rv)
    path = str(tmp_path / "r.tplkkt")
    host.save_binary(path)
    _same_system(data_loader.load_kkt_host_binary(path), host)


SPELLINGS = ["1", "2.5", "-3", "1e2", ".5", "+7.", "0", "1.e5", "1e+5", "1e-5", "-.5", "00012", "-0", "1E-400", "1e400", "-1e400",
             "4.9e-324", "123456789012345678901234567890.5", "inf", "x", " 1", "1 ", "", "nan", ".e5", "1e", "--1", "1-2", "-", ".",
             "e5", "1.5.2", "1e5e5", "1e-", "+-1", "-+1", "1_0", "0x10", "1f", "infinity", "-INF", "NaN", "nan(1)", "1e5.0", "+.5",
             "+", "+e5", "1e+", "٣", "1,5", "1e٣"]


@pytest.mark.parametrize("spelling", SPELLINGS)
def test_cost_grammar_matches_the_oracle(tmp_path, spelling):
    """every cost spelling takes the same branch in the product loader (fast decimal path or the validating one) and in the
    oracle's restatement of `<f64 as FromStr>`: the same value bit for bit, or ParseFloat with the same message"""
    dmx = _write(tmp_path, "g.dmx", "p min 2 1\na 1 2\n")
    qfc = _write(tmp_path, "g.qfc", "1\n0\n" + spelling + "\n")
    try:
        want = ("ok", orc.load_kkt_system(dmx, qfc).a.csc()[2])
    except orc.OracleError as e:
        want = ("err", e.code, str(e))
    try:
        got = ("ok", data_loader.load_kkt_host(dmx, qfc).csc()[3])
    except DataLoaderError as e:
        got = ("err", e.code, str(e))
    assert got[0] == want[0], (spelling, got, want)
    if got[0] == "ok":
        assert np.array_equal(got[1], want[1], equal_nan=True) and np.array_equal(np.signbit(got[1]), np.signbit(want[1]))
    else:
        assert got[1:] == want[1:]
