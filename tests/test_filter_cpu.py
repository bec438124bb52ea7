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


# ---- the mass-sorted layout's filter modes (pair_interact FILT == 2 / 3) ---------------------------------------------
def _i32(v):
    """two's-complement wrap to int32, as the kernel's 32-bit registers do"""
    return ((np.asarray(v, dtype=np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int64)


def _u32(v):
    return np.asarray(v, dtype=np.int64) % 2 ** 32


PRIOR_NEVER = 0xFFFFFFFF
PRIOR_NONE = 0x3FF00000 - MARGIN + 1


def thr_neg(t):
    return np.where(t == PRIOR_NEVER, 0x80000000, _u32(-t))


def test_source_has_the_sorted_filter_this_test_restates():
    assert "mrun = first ? (unsigned)(ly + (int)thi) : (unsigned)__viaddmax_s32(ly, (int)thi, (int)mrun);" in SRC
    assert "mrun = first ? ly + lgi : __viaddmax_u32(ly, lgi, mrun);" in SRC
    assert "prot = (int)mrun >= -(int)lgj;" in SRC and "prot = mrun >= thj;" in SRC
    assert "return t == kPriorNever ? 0x80000000u : 0u - t;" in SRC


def test_sorted_filter_modes_decide_like_the_plain_comparison_and_never_wrap():
    """Mode 2 keeps MINUS the target's threshold and takes a signed running maximum of L(y0) - thr_i over the eight
    pairs of a rotation, compared once with -lg_j; mode 3 an unsigned running maximum of L(y0) + lg_i compared with
    thr_j.  Both must flag a rotation exactly when some pair of it satisfies  lg + L(y0) >= thr  in plain unsigned
    arithmetic, over the whole range the operands can take (any positive binary64 for y0 and gf, every threshold
    k_pair_finish can leave, 'nothing passes')."""
    rng = np.random.default_rng(11)
    n, R = 40000, 8
    # y0 over the whole positive range incl. the extremes; gf likewise
    ly = hi(np.concatenate([10.0 ** rng.uniform(-300, 300, n - 4), [5e-324, 2.2e-308, 1.7e308, np.inf]]))[:, None] \
        + rng.integers(-3, 4, (n, R))
    ly = np.clip(ly, 0, 0x7FF00000)
    lg_j = key_lg(10.0 ** rng.uniform(-300, 300, n))
    lg_i = key_lg(10.0 ** rng.uniform(-300, 300, (n, R)))
    # thresholds: key_thr of any key minus the slack, the no-prior value, and 'never'
    khi = rng.integers(0, 0x7FF00000, (n, R))
    thr = (khi >> 2) + PRIOR_NONE - SLACK * rng.integers(0, 2, (n, R))
    thr = np.where(rng.random((n, R)) < 0.05, PRIOR_NEVER, thr)
    thr = np.where(rng.random((n, R)) < 0.05, PRIOR_NONE, thr)
    # make a good share of the rotations borderline: y0 within a few units of the threshold
    border = rng.random((n, R)) < 0.3
    ly = np.where(border & (thr != PRIOR_NEVER), np.clip(thr - lg_j[:, None] + rng.integers(-2, 3, (n, R)), 0, 0x7FF00000), ly)

    # mode 2: targets can gain a candidate.  plain rule per pair: lg_j + ly >= thr_i (unsigned, no wrap: < 2^32)
    plain2 = ((lg_j[:, None] + ly) >= thr).any(axis=1)
    neg = thr_neg(thr)
    terms = ly + _i32(neg)                      # the kernel's 32-bit signed add
    assert np.all(np.abs(terms) < 2 ** 31)      # never wraps
    mrun = terms.max(axis=1)
    mode2 = mrun >= -lg_j
    real = (thr != PRIOR_NEVER).all(axis=1)
    assert np.array_equal(mode2[real], plain2[real])
    # 'never' targets: cannot flag below y0 = 2^513 (r < 1e-154 m); a spurious flag would only cost a re-walk, a missed one a link
    assert not np.any(plain2 & ~mode2)

    # mode 3: visitors can gain a candidate.  plain rule per pair: lg_i + ly >= thr_j
    thr_j = thr[:, 0]
    plain3 = ((lg_i + ly) >= thr_j[:, None]).any(axis=1)
    sums = lg_i + ly
    assert np.all(sums < 2 ** 32)               # the unsigned add never wraps
    mode3 = sums.max(axis=1) >= thr_j
    assert np.array_equal(mode3, plain3)
    assert 0.05 < plain2.mean() < 0.95 and 0.05 < plain3.mean() < 0.95  # both outcomes are exercised


def test_sort_key_of_the_mass_sorted_layout_is_monotone_over_the_real_line():
    """k_pair_sort_keys: the IEEE bit pattern with negative values reversed orders like the values themselves."""
    assert "keys[g] = (b >> 63) ? ~b : b | 0x8000000000000000ull;" in SRC
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.normal(0, 1, 2000) * 10.0 ** rng.uniform(-200, 200, 2000), [0.0, -0.0, 1e-320, -1e-320, np.inf, -np.inf]])
    b = v.view(np.uint64)
    key = np.where(b >> np.uint64(63), ~b, b | np.uint64(0x8000000000000000))
    order = np.argsort(key, kind="stable")
    assert np.all(np.diff(v[order]) >= 0)
