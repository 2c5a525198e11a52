"""The host-side iteration schedule (SeqPlan) against the reference's stepping rule.

reference kernel.cu:129-142: the symbol position starts at 0, advances every step, wraps at the
-1 terminator, and CARRIES OVER from the settle loop into the accumulate loop.  The library
replaces that walk by: a partial period in written order, whole periods in an order rotated by
settle % len, a partial rotated period -- or, for long sequences, run-length segments.  Both
forms must generate exactly the reference's symbol stream.  CPU only (hypothesis-driven)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import lyapunov3d_b200 as lp
from lyapunov3d_b200 import api


def reference_stream(seq, n):
    body = [int(s) for s in seq[:-1]]
    return [body[i % len(body)] for i in range(n)]


def stream_from_table(d, settle, accum):
    out = list(d["sym"][:d["settle_head"]])
    out += d["rot"] * d["settle_periods"]
    out += d["rot"] * d["accum_periods"]
    out += d["rot"][:d["accum_tail"]]
    return out


def stream_from_runs(d, n):
    out = []
    while len(out) < n:
        for sym, ln in d["runs"]:
            out += [sym] * ln
    return out[:n]


seq_strings = st.text(alphabet="ABCDabcd123456789", min_size=1, max_size=24).filter(lambda s: s[0] not in "123456789" or True)


@settings(max_examples=300, deadline=None)
@given(seq_strings, st.integers(0, 300), st.integers(0, 2000))
def test_plan_generates_the_reference_symbol_stream(s, settle, accum):
    seq = lp.scene_convert_sequence(s)
    d = api.plan_describe(seq, settle, accum)
    want = reference_stream(seq, settle + accum)
    assert stream_from_runs(d, settle + accum) == want                 # run-length form, any length
    assert sum(ln for _, ln in d["runs"]) == d["len"] and all(1 <= ln <= 255 for _, ln in d["runs"])
    if d["P"] > 0:
        assert d["len"] == d["P"] <= 40
        assert stream_from_table(d, settle, accum) == want              # register-table form
    census = [want[settle:].count(k) for k in range(4)]
    assert d["cnt"] == census                                           # fast mode's analytic sum(log r)


@settings(max_examples=200, deadline=None)
@given(seq_strings)
def test_sequence_parser_matches_the_oracle(oracle, s):
    assert lp.scene_convert_sequence(s).tolist() == oracle.convert_sequence(s).tolist()


def test_period_reduction_and_limits():
    conv = lp.scene_convert_sequence
    assert api.plan_describe(conv("ABABABAB"), 3, 10)["len"] == 2
    assert api.plan_describe(conv("AAAA"), 3, 10)["sym"] == [0]
    d = api.plan_describe(conv("A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9A9B"), 5, 50)   # 281 symbols
    assert d["P"] == 0 and d["runs"] == [[0, 255], [0, 25], [1, 1]]
    long_seq = np.array([k % 3 for k in range(1024)] + [-1], np.int32)
    assert api.plan_period(long_seq, 1, 1) == 0
    too_long = np.array([k % 3 for k in range(1025)] + [-1], np.int32)
    assert api.plan_period(too_long, 1, 1) == -1
