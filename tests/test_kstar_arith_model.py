"""CPU model of the K* kernel's arithmetic (bocf_b200/csrc/kernfn.cuh: clamp_q, exp_nonpos_tab, sqrt_nonneg_root,
kern_eval_fast; posterior.cu: the digit split with the bias folded into the rounding constant).

The GPU parity tests hold the kernel to 1e-6 ... 1e-9 against the oracle; the comments in kernfn.cuh claim more (a few
ulp on exp and sqrt, "exactly 0 weight at r == 0", bit-identical digits).  This file restates those routines operation by
operation in numpy -- same constants, same order, plain multiply-add where the device fuses (<= 1 ulp per operation
apart) -- and pins the claims without a GPU.  Reference formulas: GPy/kern/src/stationary.py:153,227-234,440-444,529-533,
rbf.py:42-46, se.py:60 through oracle/kern.py.
"""
import numpy as np
import pytest

from oracle.kern import Kern

MAGIC = 6755399441055744.0                 # 1.5 * 2^52
TAB_N = 128
EXPT = (184.6649652337873, 0.005415212348111709, 1.2864023111638346e-14,
        8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01)
SCALE = {"matern52": 5.0, "matern32": 3.0, "rbf": 0.5, "se": 0.5}
GFAC = {"matern52": -5.0 / 3.0, "matern32": -3.0, "rbf": -1.0, "se": -1.0}
HIMAX = {"matern52": 0x411DE840, "matern32": 0x411DE840, "rbf": 0x4085E000, "se": 0x4085E000}
FLOOR_HI = 0x00100000                      # high word of the smallest normal double


def _hi_lo(x):
    b = np.asarray(x, dtype=np.float64).view(np.int64)
    return (b >> 32).astype(np.int64), (b & 0xFFFFFFFF).astype(np.int64)


def _from_hi_lo(hi, lo):
    return ((hi.astype(np.int64) << 32) | lo.astype(np.int64)).view(np.float64)


def clamp_q(q, kind):
    """kernfn.cuh clamp_q: signed min / max on the high word only; nz = "the clipped distance is not 0"."""
    hi, lo = _hi_lo(q)
    hi = np.where(hi >= 2**31, hi - 2**32, hi)          # signed view of the high word
    nz = hi >= FLOOR_HI
    return _from_hi_lo(np.minimum(np.maximum(hi, FLOOR_HI), HIMAX[kind]), lo), nz


def exp_tab(x, factor=1.0):
    """kernfn.cuh exp_nonpos_tab<false>: x = (128 k + j) ln2/128 + r, degree-5 polynomial, table factor * 2^(j/128)."""
    tab = factor * np.exp2(np.arange(TAB_N) / TAB_N)
    sh = x * EXPT[0] + MAGIC
    kf = sh - MAGIC
    r = -kf * EXPT[1] + x
    r = -kf * EXPT[2] + r
    n = kf.astype(np.int64)                              # the device reads it from the low word of sh
    p = EXPT[3] * r + EXPT[4]
    p = p * r + EXPT[5]
    p = p * r + 0.5
    p = p * r + 1.0
    p = p * r + 1.0
    p = p * tab[n & (TAB_N - 1)]
    return np.ldexp(p, (n >> 7).astype(np.int64))        # the device adds (n >> 7) to the exponent field


def sqrt_root(a, seed_err=0.0):
    """kernfn.cuh sqrt_nonneg_root with an rsqrt seed that is off by a relative seed_err (hardware: ~2^-22)."""
    y = (1.0 / np.sqrt(a)) * (1.0 + seed_err)
    s0 = a * y
    e = -s0 * y + 1.0
    p = 0.375 * e + 0.5
    return (s0 * e) * p + s0


def kern_eval_fast(q, nz, kind, variance):
    """kernfn.cuh kern_eval_fast: returns k and gp = (dK/dr)/r / GFAC."""
    if kind in ("se", "rbf"):
        k = exp_tab(-q, variance)
        gp = k if kind == "se" else np.where(nz, k, 0.0)
    elif kind == "matern52":
        t = sqrt_root(q)
        e = exp_tab(-t, variance)
        lin = 1.0 + t
        k = ((1.0 / 3.0) * q + lin) * e
        gp = np.where(nz, lin * e, 0.0)
    else:
        t = sqrt_root(q)
        e = exp_tab(-t, variance)
        k = (1.0 + t) * e
        gp = np.where(nz, e, 0.0)
    return k, gp


def test_model_constants_are_the_headers():
    """The numbers above are the ones compiled into the kernel (a drifted copy would test nothing)."""
    import os, re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bocf_b200", "csrc", "kernfn.cuh")).read()
    body = src[src.index("static __constant__ double EXPT[8]"):]
    nums = [float(x) for x in re.findall(r"^\s+([0-9.e+-]+),", body[:body.index("};")], flags=re.M)]
    assert tuple(nums[:6]) == EXPT
    assert "constexpr int EXP_TAB_N = %d;" % TAB_N in src
    assert "0x411DE840 : 0x4085E000" in src and "0x00100000" in src
    assert "? 5.0 : (KIND == BOCF_KERN_MATERN32) ? 3.0 : 0.5" in src
    # the cap keeps exp's argument above -700.01 with any low word under the capped high word
    for hi, arg in ((0x411DE840, np.sqrt), (0x4085E000, lambda v: v)):
        top = _from_hi_lo(np.array([hi]), np.array([0xFFFFFFFF]))[0]
        assert 700.0 <= arg(top) < 700.01


def test_clamp_q_clips_both_ends_on_the_high_word():
    tiny = np.finfo(np.float64).tiny
    q = np.array([-3.0, -1e-300, -0.0, 0.0, 5e-324, tiny / 2, tiny, 1e-300, 1.0, 1e5, 1e9, np.inf])
    for kind in SCALE:
        out, nz = clamp_q(q, kind)
        cap = 490000.0 if kind.startswith("matern") else 700.0
        assert np.all(out >= tiny) and np.all(out <= cap * (1 + 1e-6))
        # negative, zero and subnormal distances stand for r == 0; everything else is kept
        assert list(nz) == [False] * 6 + [True] * 6
        keep = (q >= tiny) & (q <= cap)
        assert np.array_equal(out[keep], q[keep])                # values in range pass through bit for bit
        assert np.all(out[~nz] < 2 * tiny)                       # "zero" is at most 4.5e-308


def test_exp_table_is_within_a_few_ulp_down_to_the_cap():
    rng = np.random.default_rng(1)
    x = -np.concatenate([rng.uniform(0, 40, 200000), rng.uniform(0, 700.0002, 200000), [0.0, 1e-300, 700.0002]])
    got, ref = exp_tab(x), np.exp(x)
    assert np.max(np.abs(got - ref) / ref) < 1e-15               # measured ~4e-16 without fused multiply-adds
    # the table can carry the signal variance: one rounding more, nothing else
    got_v = exp_tab(x, 2.75)
    assert np.max(np.abs(got_v - 2.75 * ref) / (2.75 * ref)) < 1.2e-15


@pytest.mark.parametrize("seed_err", [0.0, 2.0**-21, -2.0**-21])
def test_sqrt_root_refinement_absorbs_the_seed_error(seed_err):
    rng = np.random.default_rng(2)
    a = np.concatenate([10.0 ** rng.uniform(-30, 5.7, 100000), [np.finfo(np.float64).tiny, 490000.0]])
    got, ref = sqrt_root(a, seed_err), np.sqrt(a)
    assert np.max(np.abs(got - ref) / ref) < 4e-16


@pytest.mark.parametrize("kind", ["matern52", "matern32", "rbf"])
def test_kern_eval_fast_matches_the_reference_formulas(kind):
    rng = np.random.default_rng(3)
    variance = 1.7
    r2 = np.concatenate([10.0 ** rng.uniform(-12, 2.5, 50000), [0.0, -1e-17, -0.0, 1e-320]])   # incl. clipped distances
    q, nz = clamp_q(SCALE[kind] * r2, kind)
    k, gp = kern_eval_fast(q, nz, kind, variance)
    g = GFAC[kind] * gp
    kern = Kern(kind, 1, variance=variance, lengthscale=1.0)
    r = np.sqrt(np.clip(r2, 0, np.inf))                                   # stationary.py:153
    k_ref = kern.K_of_r(r)
    inv = 1.0 / np.where(r != 0.0, r, np.inf)                             # stationary.py:227-234
    g_ref = kern.dK_dr(r) * inv
    live = k_ref > 1e-290
    # exp(-a) carries a relative error of a * (rounding of a) on BOTH sides: the bar grows with the exponent's argument
    arg = np.sqrt(SCALE[kind] * np.clip(r2, 0, np.inf)) if kind.startswith("matern") else 0.5 * np.clip(r2, 0, np.inf)
    bar = 5e-16 * (6.0 + arg)
    assert np.all((np.abs(k - k_ref) / k_ref)[live] < bar[live])
    nzr = live & (r != 0.0) & (r2 >= np.finfo(np.float64).tiny)
    # the reference's own Matern-5/2 form (10/3 r - 5 r - ...) / r cancels at small r: allow it 2e-14 on top
    with np.errstate(invalid="ignore", divide="ignore"):
        assert np.all((np.abs(g - g_ref) / np.abs(g_ref))[nzr] < (bar + 2e-14)[nzr])
    # r == 0 (exact, clipped negative, subnormal): k is the signal variance and the weight is exactly 0
    z = ~nz
    assert z.sum() == 4 and np.all(g[z] == 0.0) and np.allclose(k[z], variance, rtol=1e-15, atol=0)


def test_se_weight_is_minus_k_everywhere():
    q, nz = clamp_q(0.5 * np.array([0.0, 1e-3, 4.0, 3000.0]), "se")
    k, gp = kern_eval_fast(q, nz, "se", 2.0)
    assert np.array_equal(GFAC["se"] * gp, -k)                             # se.py:139-147: no r == 0 exception
    assert np.allclose(k[:3], 2.0 * np.exp(-0.5 * np.array([0.0, 1e-3, 4.0])), rtol=1e-15)
    assert 0.0 < k[3] < 1e-300                                             # capped exponent: ~e^-700, never garbage


@pytest.mark.parametrize("S", [3, 4, 5, 6])
def test_digit_bias_in_the_rounding_constant_gives_the_balanced_digits(S):
    """posterior.cu / split_gemm.cu: bytes of bits(fma(x, 2^s, 1.5 * 2^52 + 0x80..80)) ^ 0x80 == balanced digits of rint(x 2^s)."""
    rng = np.random.default_rng(4)
    bias = int("80" * S, 16)
    scale = 2.0 ** (8 * S - 2)
    x = np.concatenate([rng.uniform(-1, 1, 100000), [0.0, 1.0 - 2.0**-40, -1.0 + 2.0**-40, 0.5 * 2.0**-(8 * S - 2), 1.5 * 2.0**-(8 * S - 2)]])
    x = x[np.abs(x) * scale < 2.0 ** (8 * S - 2)]
    prod = x * scale                                          # power-of-two scale: exact, so the add alone rounds (like the fma)
    bits = (prod + (MAGIC + float(bias))).view(np.int64)
    Y = np.rint(prod).astype(np.int64)                        # ties to even, like the add onto the (even) constant
    for t in range(S):
        byte = ((bits >> (8 * t)) & 0xFF) ^ 0x80
        got = np.where(byte >= 128, byte - 256, byte)         # two's-complement int8, as the tensor core reads it
        d = ((Y + 128) & 255) - 128
        assert np.array_equal(got, d), t
        Y = (Y - d) >> 8
    assert np.all(Y == 0)
