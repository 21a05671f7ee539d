"""CPU checks (numpy) of the arithmetic identities the CUDA path relies on and DESIGN.md quotes -- the formulas, not the kernels:
split-f16 operands (precision="split"), the fused-LayerNorm statistics merge, the GELU(erf) approximation of the GEMM epilogue."""
import numpy as np
from scipy.special import erf

F32 = np.float32


def _split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(F32)).astype(np.float16)
    return hi, lo


def test_split_f16_three_products_recover_f32_level_accuracy():
    """include/udb.h udb_gemm_t.a_split_k: A = [hi | lo], W = [W_hi | W_hi | W_lo] accumulates hi*W_hi + lo*W_hi + hi*W_lo in f32
    (error 2^-21 instead of 2^-11)."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((64, 1024)).astype(F32)
    w = (rng.standard_normal((96, 1024)) / 32).astype(F32)
    exact = a.astype(np.float64) @ w.astype(np.float64).T
    ah, al = _split(a)
    wh, wl = _split(w)
    f = lambda x: x.astype(F32)
    plain = f(ah) @ f(wh).T
    split = f(ah) @ f(wh).T + f(al) @ f(wh).T + f(ah) @ f(wl).T
    scale = np.abs(exact).mean()
    e_plain = np.abs(plain - exact).mean() / scale
    e_split = np.abs(split - exact).mean() / scale
    assert 5e-5 < e_plain < 1e-3          # f16 operand rounding
    assert e_split < 2e-6                 # the dropped lo*lo term and f32 summation
    assert e_split < e_plain / 100


def test_layernorm_partial_statistics_merge():
    """Fused LayerNorm (udb_gemm_t.ln_*): producers emit per-part {mean, centred sum of squares} computed against a pivot (the
    part's first value), the consumer merges equal-sized parts: mean = mean of means, M2 = sum M2_p + n_p sum (mean_p - mean)^2.
    No E[x^2] - E[x]^2 cancellation even when |mean| >> std."""
    rng = np.random.default_rng(1)
    for offset in (0.0, 50.0, 3000.0):
        x = (rng.standard_normal((8, 1024)) + offset).astype(F32)
        parts = x.reshape(8, 8, 128)
        pivot = parts[:, :, :1]
        d = (parts - pivot).astype(F32)
        s1 = d.sum(-1, dtype=F32)
        s2 = (d * d).sum(-1, dtype=F32)
        mean_p = (pivot[..., 0] + s1 / F32(128)).astype(F32)
        m2_p = (s2 - s1 * s1 / F32(128)).astype(F32)
        mean = mean_p.mean(-1, dtype=F32)
        m2 = (m2_p.sum(-1, dtype=F32) + F32(128) * ((mean_p - mean[:, None]) ** 2).sum(-1, dtype=F32)).astype(F32)
        var = m2 / F32(1024)
        ref_var = x.astype(np.float64).var(-1)
        ref_mean = x.astype(np.float64).mean(-1)
        assert np.abs(mean - ref_mean).max() < 1e-6 * max(1.0, offset) + 1e-6
        assert np.abs(var / ref_var - 1).max() < 2e-5, (offset, np.abs(var / ref_var - 1).max())
        naive = (x * x).mean(-1, dtype=F32) - x.mean(-1, dtype=F32) ** 2          # what the merge formula avoids
        if offset >= 3000.0:
            assert np.abs(naive / ref_var - 1).max() > 1e-2


def test_gelu_erf_approximation_of_the_gemm_epilogue():
    """ptx.cuh gelu_erf_pair: erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16 (Abramowitz & Stegun 7.1.28, |error| <= 3e-7)."""
    a = [0.0705230784, 0.0422820123, 0.0092705272, 0.0001520143, 0.0002765672, 0.0000430638]
    x = np.linspace(-9, 9, 200001)
    z = np.abs(x) / np.sqrt(2.0)
    q = np.ones_like(z)
    for i, c in enumerate(a):
        q = q + c * z ** (i + 1)
    r = q ** -16.0
    gelu = 0.5 * x * (1.0 + np.copysign(1.0 - r, x))
    ref = 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))
    assert np.abs(gelu - ref).max() < 2e-6      # absolute; the epilogue's output is rounded to f16 (2^-11 relative) afterwards
