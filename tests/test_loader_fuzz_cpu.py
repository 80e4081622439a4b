"""Differential fuzzing of the product loader (tpl_load_kkt, tpl_loader.cpp) against the oracle's independent restatement
of src/utils/data_loader.rs (oracle/lanczos_oracle.cpp): on token soups that look like `.dmx` / `.qfc` files both must take
the same branch -- the same DataLoaderError code and message, or the same matrix entry for entry."""
import os
import tempfile

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as orc
from two_pass_lanczos_b200 import data_loader
from two_pass_lanczos_b200.error import DataLoaderError

TOKENS = ["c", "p", "a", "n", "min", "max", "0", "1", "2", "3", "4", "7", "+2", "-1", "x", "1.5", "1e3", ".5", "2.", "inf", "-inf",
          "nan", "1e400", "0x1", "١", "aa", "P", "A"]
SEPS = [" ", "  ", "\t", " \t "]
line = st.lists(st.sampled_from(TOKENS), min_size=0, max_size=6).flatmap(
    lambda toks: st.lists(st.sampled_from(SEPS), min_size=len(toks) + 1, max_size=len(toks) + 1).map(
        lambda seps: "".join(s + t for s, t in zip(seps, toks)) + (seps[-1] if len(seps[-1]) > 1 else "")))
eol = st.sampled_from(["\n", "\n", "\n", "\r\n"])
text = st.lists(st.tuples(line, eol), min_size=0, max_size=12).map(lambda ls: "".join(a + b for a, b in ls))
# files that are mostly well formed (with an occasional defect), so that the deep branches and the success path are reached
@st.composite
def nearly_good_pair(draw):
    nodes = draw(st.integers(1, 5))
    arcs = draw(st.integers(0, 6))
    n_lines = arcs + draw(st.sampled_from([0, 0, 0, 0, -1, 1]))
    idx = st.integers(1, nodes) if draw(st.integers(0, 5)) else st.integers(0, nodes + 1)
    body = []
    for _ in range(max(n_lines, 0)):
        u, v = draw(idx), draw(idx)
        body.append(f"a {u} {v}" + draw(st.sampled_from(["", " 0 9 9", "\t7"])))
        if not draw(st.integers(0, 7)):
            body.append(draw(line))
    dmx = draw(st.sampled_from(["", "c generated\n"])) + f"p min {nodes} {arcs}\n" + "".join(x + "\n" for x in body)
    m_line = arcs if draw(st.integers(0, 7)) else arcs + 1
    good_num = st.sampled_from(["1", "2.5", "-3", "1e2", ".5", "+7.", "0", "1.e5", "1e+5", "1e-5", "-.5", "00012", "-0", "1E-400",
                                "1e400", "-1e400", "4.9e-324", "123456789012345678901234567890.5"])
    num = good_num if draw(st.integers(0, 5)) else st.sampled_from(
        ["inf", "x", " 1", "1 ", "", "nan", ".e5", "1e", "--1", "1-2", "-", ".", "e5", "1.5.2", "1e5e5", "1e-", "+-1", "-+1", "1_0",
         "0x10", "1f", "infinity", "-INF", "NaN", "nan(1)", "1e5.0"])
    n_costs = draw(st.sampled_from([arcs, arcs, arcs, 0, max(arcs - 1, 0), arcs + 2]))
    qfc = f"{m_line}\n" + "".join("skipped\n" for _ in range(arcs)) + "".join(draw(num) + "\n" for _ in range(n_costs))
    return dmx, qfc


def _both(dmx_text, qfc_text):
    with tempfile.TemporaryDirectory() as d:
        dmx, qfc = os.path.join(d, "f.dmx"), os.path.join(d, "f.qfc")
        with open(dmx, "w", encoding="utf-8", newline="") as f:
            f.write(dmx_text)
        with open(qfc, "w", encoding="utf-8", newline="") as f:
            f.write(qfc_text)
        try:
            ref = orc.load_kkt_system(dmx, qfc)
            ref_out = ("ok", ref.num_nodes, ref.num_arcs, ref.a.csc())
        except orc.OracleError as e:
            ref_out = ("err", e.code, str(e))
        try:
            host = data_loader.load_kkt_host(dmx, qfc)
            got = ("ok", host.num_nodes, host.num_arcs, host.csc()[1:])
        except DataLoaderError as e:
            got = ("err", e.code, str(e))
    return got, ref_out


def _same(got, ref):
    assert got[0] == ref[0], (got[:3], ref[:3])
    if got[0] == "err":
        if got[1] == 109:  # a malformed `a` line: the reference panics there, the oracle reports the same code
            assert ref[1] == 109
        else:
            assert got[1:] == ref[1:], (got, ref)
        return
    assert got[1:3] == ref[1:3]
    for x, y in zip(got[3], ref[3]):
        assert np.array_equal(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), equal_nan=True)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(text, text)
def test_token_soup(dmx_text, qfc_text):
    _same(*_both(dmx_text, qfc_text))


@settings(max_examples=250, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(nearly_good_pair())
def test_nearly_well_formed_files(pair):
    _same(*_both(*pair))
