"""Pins oracle/native_oracle.c to the reference's own known answers.

Vectors restate native-helper/src/lib.rs:683-1173 (cargo tests) and
native-helper/tests/test_python_bindings.py (scipy / pyloudnorm comparisons).
"""
import numpy as np
import pytest

from oracle import native


def fp(data, **kw):
    return native.find_peaks(np.asarray(data, dtype=np.float32), **kw)[0].tolist()


def test_local_maxima_vectors():                         # lib.rs:683-717
    assert native.local_maxima(np.array([0, 1, 0, 2, 0], np.float32)).tolist() == [1, 3]
    for short in ([], [1.0], [1.0, 2.0]):
        assert native.local_maxima(np.array(short, np.float32)).tolist() == []
    assert native.local_maxima(np.array([0, 1, 1, 0], np.float32)).tolist() == [1]
    assert native.local_maxima(np.array([0, 1, 1, 1, 0], np.float32)).tolist() == [2]
    assert native.local_maxima(np.array([1, 2, 3, 4, 5], np.float32)).tolist() == []
    assert native.local_maxima(np.array([5, 4, 3, 2, 1], np.float32)).tolist() == []
    assert fp([1, 1, 1, 1]) == []                        # lib.rs:841-850


def test_height_and_distance_vectors():                  # lib.rs:721-755, 804-814
    assert fp([0, 1, 0, 2, 0], height=1.5) == [3]
    assert fp([0, 3, 0, 5, 0], distance=3) == [3]
    assert fp([0, 3, 0, 0, 0, 5, 0], distance=3) == [1, 5]
    assert fp([0, 2, 0, 1, 0, 3, 0], height=1.5, distance=3) == [1, 5]


def test_distance_tie_prefers_lower_index():             # lib.rs:446-451
    assert fp([0, 5, 0, 5, 0], distance=3) == [1]


def test_prominence_vectors():                           # lib.rs:778-800, 816-839
    assert fp([0, 1, 0.5, 2, 0], prominence=1.0) == [3]
    assert fp([0, 5, 0, 5, 0], prominence=4.0) == [1, 3]
    assert fp([0, 5, 4.5, 5, 0, 3, 0], height=2.0, distance=2, prominence=1.0) == [1, 3, 5]


def test_find_peaks_matches_scipy():                     # test_python_bindings.py:119-145
    from scipy.signal import find_peaks as sfp
    rng = np.random.default_rng(42)
    x = np.linspace(0, 10 * np.pi, 500).astype(np.float32)
    d = np.abs((np.sin(x) + 0.3 * rng.standard_normal(500)).astype(np.float32))
    d /= np.max(d)
    assert fp(d, height=0.25, distance=20) == sfp(d, height=0.25, distance=20)[0].tolist()
    rng = np.random.default_rng(123)
    d = np.abs(rng.standard_normal(200).astype(np.float32))
    d /= np.max(d)
    assert fp(d, prominence=0.05) == sfp(d, prominence=0.05)[0].tolist()
    for seed in range(5):
        d = np.abs(np.random.default_rng(seed).standard_normal(5000)).astype(np.float32)
        for dist in (1, 7, 100, 999):
            assert fp(d, height=0.5, distance=dist) == sfp(d, height=0.5, distance=dist)[0].tolist()


def test_window_max_vectors():                           # lib.rs:892-966
    r = native.resample_preserve_maxima
    a = lambda v: np.array(v, np.float32)
    assert r(a([1, 3, 2, 4]), 4).tolist() == [1, 3, 2, 4]
    assert r(a([1, 5, 2, 4, 3, 6]), 3).tolist() == [5, 4, 6]
    assert r(a([1, 2, 3]), 5).tolist() == [1, 1, 2, 2, 3]
    assert r(a([7]), 4).tolist() == [7, 7, 7, 7]
    assert r(a([1, 5]), 6).tolist() == [1, 1, 1, 5, 5, 5]
    assert r(a([3, 1, 4, 1, 5]), 20).tolist() == [3] * 4 + [1] * 4 + [4] * 4 + [1] * 4 + [5] * 4
    assert r(a([2, 8, 3, 7, 1]), 5).tolist() == [2, 8, 3, 7, 1]
    assert r(np.array([1, 5, 2, 4], np.float64), 2).tolist() == [5, 4]
    with pytest.raises(ValueError):
        r(a([1, 2]), 0)


def test_resample_kats():                                # lib.rs:855-890, bindings tests :154-189
    from scipy.signal import resample as scipy_resample
    r = native.resample
    out = r(np.array([1, 2, 3, 4], np.float32), 4)
    assert out.dtype == np.float32 and np.allclose(out, [1, 2, 3, 4], atol=1e-5)
    assert r(np.zeros(0, np.float32), 0).size == 0
    assert r(np.zeros(0, np.float32), 5).tolist() == [0.0] * 5
    assert r(np.array([1, 2], np.float32), 0).size == 0
    sine = np.sin(2.0 * np.float32(np.pi) * np.arange(8, dtype=np.float32) / np.float32(8)).astype(np.float32)
    out = r(sine, 4)
    assert out.size == 4 and abs(out[0]) < 0.1 and abs(out[1] - 1.0) < 0.1
    assert r(np.array([0, 1, 0], np.float32), 6).size == 6
    data = np.random.default_rng(99).standard_normal(160).astype(np.float32)
    np.testing.assert_allclose(r(data, 80), scipy_resample(data, 80).astype(np.float32), atol=0.2)
    data = np.array([0, 1, 0, -1, 0], np.float32)
    np.testing.assert_allclose(r(data, 10), scipy_resample(data, 10).astype(np.float32), atol=1e-4)
    out = r(np.array([1, 2, 3, 4], np.float64), 2)
    assert out.dtype == np.float32 and out.size == 2
    # the product's host-side clip loader (audio_utils.resample_fft) follows the same restatement
    from audio_pattern_detector_b200.audio_utils import resample_fft
    x = np.random.default_rng(5).standard_normal(1001).astype(np.float32)
    for m in (500, 1000, 1002, 2003):
        np.testing.assert_allclose(resample_fft(x, m), r(x, m), rtol=0, atol=1e-6)


def test_k_weighting_and_loudness():                     # lib.rs:1015-1103, bindings :253-300
    bs, as_, bh, ah = native.k_weighting_coefficients(8000.0)
    assert abs(bs[0] - 1.32773315) < 1e-5 and as_[0] == 1.0
    assert abs(bh[0] - 0.97080775) < 1e-5 and ah[0] == 1.0
    assert native.integrated_loudness(np.zeros(8000, np.float32), 8000) == float("-inf")
    t = np.arange(8000, dtype=np.float32) / 8000
    assert -10.0 < native.integrated_loudness(np.sin(2 * np.pi * 1000 * t).astype(np.float32), 8000) < 0.0
    rng = np.random.default_rng(42)
    d = (rng.standard_normal(16000) * 0.3).astype(np.float32)
    assert abs(native.integrated_loudness(d, 8000) - (-8.438312960262843)) < 0.05
    t = np.arange(2400, dtype=np.float32) / 8000
    assert np.isfinite(native.integrated_loudness(np.sin(2 * np.pi * 440 * t).astype(np.float32), 8000, 0.3))
    out = native.loudness_normalize(np.array([0.5, -0.5, 0.8, -0.8], np.float32), -60.0, -20.0)
    assert out.dtype == np.float32 and np.all(np.abs(out) <= 1.0)
    out = native.loudness_normalize(np.array([0.1, -0.1], np.float32), -22.0, -16.0)
    assert abs(out[0] - 0.1 * 10 ** (6 / 20)) < 1e-4
    # silence: gain = +inf, 0*inf = NaN survives the clamp (lib.rs:225), non-zero saturates
    out = native.loudness_normalize(np.array([0.0, 1e-9, -1e-9], np.float32), float("-inf"), -16.0)
    assert np.isnan(out[0]) and out[1] == 1.0 and out[2] == -1.0


def test_loudness_matches_independent_scipy_formulation():
    """Same BS.1770 gating written a second way (scipy lfilter + cumsum)."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(7)
    for n, sr in ((48000, 8000), (100000, 16000), (3000, 8000)):
        x = (rng.standard_normal(n) * 0.1).astype(np.float32)
        x[n // 3:n // 2] *= 0.001
        bs, as_, bh, ah = native.k_weighting_coefficients(sr)
        y = lfilter(bh, ah, lfilter(bs, as_, x.astype(np.float64)))
        P = np.concatenate([[0.0], np.cumsum(y * y)])
        T = n / sr
        blk = T if T < 0.5 else 0.4
        nb = int(np.floor((T - blk) / (blk * 0.25) + 0.5)) + 1
        ms = []
        for j in range(nb):
            lo, hi = int(j * blk * sr * 0.25), min(int(j * blk * sr * 0.25 + blk * sr), n)
            if lo < hi:
                ms.append((P[hi] - P[lo]) / (hi - lo))
        ms = np.array(ms)
        l = -0.691 + 10 * np.log10(ms)
        g = -0.691 + 10 * np.log10(ms[l >= -70].mean()) - 10
        want = -0.691 + 10 * np.log10(ms[(l > g) & (l >= -70)].mean())
        assert abs(native.integrated_loudness(x, sr, blk) - want) < 1e-9


def test_pearson_vectors():                              # lib.rs:1107-1173
    p = native.pearson_correlation
    a = np.array([1, 2, 3, 4, 5], np.float32)
    assert abs(p(a, a) - 1) < 1e-12 and abs(p(a, -a) + 1) < 1e-12
    assert p(np.full(5, 3, np.float32), a) == 0.0
    assert p(np.array([], np.float32), np.array([], np.float32)) == 0.0
    assert abs(p(a, 3 * a + 10) - 1) < 1e-12
    with pytest.raises(ValueError):
        p(a, a[:3].copy())
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(300).astype(np.float32), rng.standard_normal(300).astype(np.float32)
    assert abs(p(x, y) - np.corrcoef(x.astype(np.float64), y.astype(np.float64))[0, 1]) < 1e-12
