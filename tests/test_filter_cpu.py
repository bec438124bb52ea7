"""The integer candidate filter of the pair-shared kernel's most-attracting tracking (essa_b200/csrc/force_paired.cu:
key_lg / key_thr / pack_key), restated in numpy from the constants in the source, against the property the kernel
relies on:  a partner whose exact key beats (or ties) the best key so far ALWAYS passes the filter,

    key(M_c) >= key(M_b)   ==>   (hi(gf_c) >> 2) + hi(y0_c)  >=  key_thr(key(M_b)),

for metrics M = gf * |d|^-3 * y0 formed from a 21-bit rsqrt seed y0 whose refinement |d|^-3 = y0^3 (1 + eps),
|eps| <= 3e-6.  (Object.cpp:84-94 ranks partners by |attraction|^2 / gf = gf / r^4.)  No GPU needed."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "essa_b200", "csrc", "force_paired.cu")).read()


def _const(name):
    m = re.search(r"constexpr\s+\w+(?:\s+\w+)*\s+%s\s*=\s*([^;]+);" % name, SRC)
    assert m, name
    expr = m.group(1).split("//")[0]
    expr = re.sub(r"(\d)u(ll)?\b", r"\1", expr)
    return expr.strip()


IDX_BITS = int(_const("kKeyIdxBits"))
MARGIN = int(eval(_const("kKeyMargin")))
SLACK = int(eval(_const("kPriorSlack")))
IDX_MASK = (1 << IDX_BITS) - 1


def hi(v):
    return (np.asarray(v, dtype=np.float64).view(np.uint64) >> np.uint64(32)).astype(np.int64)


def key_lg(gf):
    return hi(gf) >> 2


def pack_key(m, idx):
    b = np.asarray(m, dtype=np.float64).view(np.uint64).astype(object)
    return np.array([(int(x) & ~IDX_MASK) | (IDX_MASK - int(i)) for x, i in zip(b, idx)], dtype=object)


def key_thr(key):
    return np.array([((int(k) >> 32) >> 2) + (0x3FF00000 - MARGIN + 1) for k in key], dtype=np.int64)


def seed21(y):
    """keep the high word only, like MUFU.RSQ64H + the zeroed low word"""
    b = np.asarray(y, dtype=np.float64).view(np.uint64) & np.uint64(0xFFFFFFFF00000000)
    return b.view(np.float64)


def test_source_constants_are_the_ones_this_test_assumes():
    assert "return (unsigned)__double2hiint(gf) >> 2;" in SRC
    assert "((unsigned)(key >> 32) >> 2) + (0x3FF00000u - kKeyMargin + 1u)" in SRC
    assert MARGIN > 0.0861 * 1.25 * 2 ** 20 + 8      # delta = f - log2(1 + f) in [-0.0861, 0], once /4 and once whole
    assert 0 < SLACK < 0x3FF00000 - MARGIN


def test_piecewise_linear_log_error_bound():
    f = np.linspace(0.0, 1.0, 200001)[:-1]
    delta = f - np.log2(1.0 + f)
    assert delta.max() <= 1e-15 and delta.min() >= -0.08608


def _metric(gf, y0, eps):
    u = (y0 * y0) * y0 * (1.0 + eps)
    return (gf * u) * y0


def test_a_better_or_equal_partner_always_passes():
    rng = np.random.default_rng(7)
    n = 60000
    gf_b = 10.0 ** rng.uniform(-3, 30, n)
    gf_c = 10.0 ** rng.uniform(-3, 30, n)
    y_b = seed21(10.0 ** rng.uniform(-14, -2, n))
    eps_b = rng.uniform(-3e-6, 3e-6, n)
    eps_c = rng.uniform(-3e-6, 3e-6, n)
    m_b = _metric(gf_b, y_b, eps_b)
    # candidates within a few 21-bit ulps of a tie with the best (the hard cases), on both sides of it
    y_tie = (m_b / gf_c) ** 0.25
    steps = rng.integers(-3, 4, n)
    y_c = seed21(y_tie * (1.0 + steps * 2.0 ** -20))
    m_c = _metric(gf_c, y_c, eps_c)
    idx_b = rng.integers(0, 1 << IDX_BITS, n)
    idx_c = rng.integers(0, 1 << IDX_BITS, n)
    k_b, k_c = pack_key(m_b, idx_b), pack_key(m_c, idx_c)
    beats = np.array([c >= b for b, c in zip(k_b, k_c)])
    assert 0.2 < beats.mean() < 0.8  # both outcomes are exercised
    a_c = key_lg(gf_c) + hi(y_c)
    thr = key_thr(k_b)
    assert np.all(a_c[beats] >= thr[beats])
    # ... and far-away winners too (orders of magnitude stronger)
    y_far = seed21(y_tie * 10.0 ** rng.uniform(0, 2, n))
    a_far = key_lg(gf_c) + hi(y_far)
    assert np.all(a_far >= thr)
    # every candidate passes the threshold of its own key: what k_pair_finish's acceptance rule builds on
    assert np.all(a_c >= key_thr(k_c))


def test_the_filter_rejects_what_is_clearly_weaker():
    """Not needed for correctness, but it is the point of the filter: a partner 4x weaker than the best is skipped."""
    rng = np.random.default_rng(8)
    n = 20000
    gf_b = 10.0 ** rng.uniform(5, 25, n)
    gf_c = 10.0 ** rng.uniform(5, 25, n)
    y_b = seed21(10.0 ** rng.uniform(-13, -5, n))
    m_b = _metric(gf_b, y_b, 0.0)
    y_c = seed21((m_b / 4.0 / gf_c) ** 0.25)
    k_b = pack_key(m_b, np.zeros(n, dtype=np.int64))
    a_c = key_lg(gf_c) + hi(y_c)
    assert np.all(a_c < key_thr(k_b))
