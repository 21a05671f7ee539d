"""CPU restatement of the ALGORITHM of unidepth_b200/csrc/attention.cu (attn_fwd2_kernel) in numpy float32 / float16 -- not the
CUDA code, the arithmetic it performs -- checked against an exact float64 softmax(QK^T)V:

  * two independent online-softmax streams per query row (the 64-key halves of every 128-key tile), chunks of 32 keys;
  * speculation: probabilities of a chunk are computed against the CURRENT reference; the reference is kept while their sum
    stays <= 2^12 (then no probability can exceed the f16 range), otherwise the chunk is redone with its true maximum and the
    accumulator / row sum are rescaled only if the maximum outgrew the reference by more than 2^8 (lazy rescale);
  * P rounded to f16 before P.V, row sums from the unrounded f32 probabilities, f32 accumulation;
  * every 4th pair of exp2 by the degree-4 polynomial + exponent-field trick of exp2_poly_pair;
  * merge of the two streams: O = (a0 O0 + a1 O1) / (a0 l0 + a1 l1), a_g = 2^(m_g - max m), a stream without keys drops out.

What it pins down on a CPU-only machine: the bounds the kernel's comments claim (polynomial error, no f16 overflow below the
speculation limit, exactness of the merge), and the error level the GPU tests assert (tests/test_attention_gpu.py)."""
import numpy as np

F32 = np.float32
LOG2E = 1.4426950408889634
SPEC_LIMIT = F32(4096.0)
RESCALE = F32(8.0)
POLY = [F32(0.9999992847442627), F32(0.6931217908859253), F32(0.240247443318367), F32(0.05591785907745361),
        F32(0.009570102207362652)]


def exp2_poly(x):
    """exp2_poly_pair: x = k + f with k = round(x) by the 1.5 * 2^23 trick, 2^f by Horner, 2^k added into the exponent field."""
    x = np.maximum(x.astype(F32), F32(-100.0))
    magic = F32(12582912.0)
    t = (x + magic).astype(F32)
    n = (t - magic).astype(F32)
    f = (x - n).astype(F32)
    q = np.full_like(f, POLY[4])
    for c in (POLY[3], POLY[2], POLY[1], POLY[0]):
        q = (q * f + c).astype(F32)
    bits = q.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))
    return bits.view(F32)


def exp2_mixed(x):
    """The kernel's mix inside a 16-score half chunk: pairs 3 and 7 of 8 take the polynomial, the rest the hardware ex2."""
    out = np.exp2(x.astype(np.float64)).astype(F32)
    pairs = (np.arange(x.shape[-1]) >> 1) % 4 == 3
    out[..., pairs] = exp2_poly(x[..., pairs])
    return out


def stream_attention(s, v, counters):
    """One stream of one query row block: s [rows, keys] scaled scores (log2 domain) of the keys this stream owns, in order;
    v [keys, d].  Returns (m_ref, l, O) per row."""
    rows = s.shape[0]
    m_ref = np.full(rows, -np.inf, F32)
    l = np.zeros(rows, F32)
    O = np.zeros((rows, v.shape[1]), F32)
    for c0 in range(0, s.shape[1], 32):
        sc = s[:, c0:c0 + 32]
        vc = v[c0:c0 + 32].astype(F32)
        first = c0 == 0
        careful = np.full(rows, first)
        if not first:
            with np.errstate(over="ignore", invalid="ignore"):
                e = np.concatenate([exp2_mixed(sc[:, i:i + 16] - m_ref[:, None]) for i in range(0, sc.shape[1], 16)], 1)
            psum = e.sum(1, dtype=F32)
            careful = ~(psum <= SPEC_LIMIT)
            # the kernel votes per warp (32 rows); per-row here is the same arithmetic for the rows that pass
            assert np.all(e[~careful] <= SPEC_LIMIT), "a kept probability exceeds the speculation limit"
            counters["kept"] += int((~careful).sum())
        alpha = np.ones(rows, F32)
        if careful.any():
            counters["first" if first else "redone"] += int(careful.sum())
            m_chunk = sc.max(1)
            need = careful & (m_chunk > m_ref + RESCALE)
            alpha[need] = np.exp2((m_ref[need] - m_chunk[need]).astype(np.float64)).astype(F32)
            m_ref = np.where(need, m_chunk, m_ref).astype(F32)
            e2 = np.concatenate([exp2_mixed(sc[:, i:i + 16] - m_ref[:, None]) for i in range(0, sc.shape[1], 16)], 1)
            if first:
                e, psum = e2, e2.sum(1, dtype=F32)
            else:
                e[careful] = e2[careful]
                psum = np.where(careful, e2.sum(1, dtype=F32), psum)
        O = O * alpha[:, None]
        l = (l * alpha + psum).astype(F32)
        p16 = e.astype(np.float16)
        assert np.isfinite(p16).all(), "P overflowed the f16 range"
        O = (O + p16.astype(F32) @ vc).astype(F32)
    return m_ref, l, O


def two_stream_attention(q, k, v, scale):
    counters = {"kept": 0, "first": 0, "redone": 0}
    s = (q.astype(F32) @ k.astype(F32).T * F32(scale * LOG2E)).astype(F32)
    n = k.shape[0]
    idx = np.arange(n)
    own = [idx[(idx % 128) // 64 == g] for g in (0, 1)]
    res = []
    for g in (0, 1):
        if len(own[g]) == 0:
            res.append((np.full(q.shape[0], -np.inf, F32), np.zeros(q.shape[0], F32), np.zeros((q.shape[0], v.shape[1]), F32)))
        else:
            res.append(stream_attention(s[:, own[g]], v[own[g]], counters))
    (m0, l0, O0), (m1, l1, O1) = res
    m0 = np.where(l0 > 0, m0, -np.inf)
    m1 = np.where(l1 > 0, m1, -np.inf)
    m = np.maximum(m0, m1)
    a0 = np.exp2((m0 - m).astype(np.float64)).astype(F32)
    a1 = np.exp2((m1 - m).astype(np.float64)).astype(F32)
    inv = F32(1.0) / (l0 * a0 + l1 * a1)
    out = O0 * (a0 * inv)[:, None] + np.where((l1 > 0)[:, None], O1 * (a1 * inv)[:, None], 0)
    return out.astype(np.float16), counters


def exact(q, k, v, scale):
    s = q.astype(np.float64) @ k.astype(np.float64).T * scale
    p = np.exp(s - s.max(1, keepdims=True))
    return (p / p.sum(1, keepdims=True)) @ v.astype(np.float64)


def test_polynomial_exp2_error_and_range():
    x = np.linspace(-100.0, 12.5, 400001).astype(F32)
    y = exp2_poly(x).astype(np.float64)
    ref = np.exp2(x.astype(np.float64))
    assert np.max(np.abs(y / ref - 1.0)) < 3.5e-6            # attention.cu: "relative error 2.7e-6" + f32 evaluation
    assert np.all(exp2_poly(np.array([-1e4, -150.0], F32)) < 1e-29)   # clamped: 0 in f16 either way


def test_two_streams_match_exact_softmax_on_gaussian_inputs():
    rng = np.random.default_rng(0)
    for sq, sk in ((70, 300), (64, 1611), (33, 40), (20, 65), (16, 129)):
        q, k, v = (rng.standard_normal((n, 64)).astype(np.float16) for n in (sq, sk, sk))
        out, cnt = two_stream_attention(q, k, v, 64 ** -0.5)
        err = np.abs(out.astype(np.float64) - exact(q, k, v, 64 ** -0.5)).max()
        assert err < 4e-3, (sq, sk, err)                      # the bound tests/test_attention_gpu.py asserts for the kernel
        assert cnt["redone"] <= 0.02 * cnt["kept"] + 2, cnt   # after a stream's first chunk the speculation almost always holds


def test_growing_logits_exercise_the_redo_and_rescale_paths():
    rng = np.random.default_rng(1)
    sq, sk = 48, 900
    q = (rng.standard_normal((sq, 64)) * 2.0).astype(np.float16)
    k = (rng.standard_normal((sk, 64)) * (1.0 + 5.0 * np.arange(sk)[:, None] / sk)).astype(np.float16)
    v = rng.standard_normal((sk, 64)).astype(np.float16)
    out, cnt = two_stream_attention(q, k, v, 64 ** -0.5)
    assert cnt["redone"] > 2 * sq, cnt                        # the maxima keep growing: chunks are redone and O / l rescaled
    ref = exact(q, k, v, 64 ** -0.5)
    assert np.isfinite(out).all()
    assert np.abs(out.astype(np.float64) - ref).max() < 8e-3   # tests/test_attention_gpu.py's bound for this kind of input


def test_reference_may_lag_by_up_to_the_speculation_limit_without_overflow():
    # one key per chunk sits exactly 11.5 (log2) above the reference set by the first chunk: kept by the sum check, P = 2^11.5 in f16
    sq, sk = 8, 256
    q = np.zeros((sq, 64), np.float16)
    q[:, 0] = 8.0
    k = np.zeros((sk, 64), np.float16)
    k[40::32, 0] = np.float16(11.5 / LOG2E)                    # score * log2(e) = 11.5 with scale 1/8 and q0 = 8
    v = np.random.default_rng(2).standard_normal((sk, 64)).astype(np.float16)
    out, cnt = two_stream_attention(q, k, v, 64 ** -0.5)
    assert cnt["kept"] > 0
    assert np.abs(out.astype(np.float64) - exact(q, k, v, 64 ** -0.5)).max() < 4e-3
