"""CPU tests of oracle/metrics.py (numpy restatement of the restatable acceptance metrics of
evaluation_metrics/calculate_intrusive_se_metrics.py:37-48,90-109): known answers and invariances."""
import numpy as np

from oracle import metrics as M


def _speechlike(n, fs, seed):
    from urgent2026_challenge_track1_b200.synth import synth_pair
    c, x = synth_pair(1, n, fs, seed=seed)
    rng = np.random.RandomState(seed)
    breath = rng.randn(n) * float(c.std()) * 10 ** (-25 / 20)          # broadband floor: a well-conditioned reference
    return c[0].numpy().astype(np.float64) + breath, x[0].numpy().astype(np.float64) + breath


def test_estoi_known_answers():
    clean, noisy = _speechlike(16000 * 3, 16000, 3)
    assert abs(M.estoi(clean, clean, 16000) - 1.0) < 1e-9
    e_noisy = M.estoi(clean, noisy, 16000)
    e_worse = M.estoi(clean, clean + 4 * (noisy - clean), 16000)
    assert 0.0 < e_worse < e_noisy < 1.0
    assert abs(M.estoi(clean, 0.5 * noisy, 16000) - e_noisy) < 1e-6      # scale invariant
    assert abs(M.estoi(clean, noisy, 16000) - e_noisy) < 1e-12          # deterministic (seeded like the reference)
    assert M.estoi(clean[:2000], noisy[:2000], 16000) == 1e-5           # too short: pystoi's fallback value
    obm, cf = M.thirdoct(10000, 512, 15, 150)
    assert obm.shape == (15, 257) and abs(cf[-1] - 150 * 2 ** (14 / 3)) < 1e-9 and obm.sum(1).min() >= 1


def test_estoi_resamples_like_octave():
    h = M._resample_window_oct(10000, 48000)
    assert h.size % 2 == 1 and abs(h.sum() / 5 - 1.0) < 0.05             # p = 5 after gcd: unit passband gain per phase
    x = np.sin(2 * np.pi * 440 * np.arange(48000) / 48000)
    y = M.resample_oct(x, 10000, 48000)
    assert y.size == 10000 and abs(np.abs(y[1000:9000]).max() - 1.0) < 0.02


def test_sdr_known_answers():
    rng = np.random.RandomState(0)
    x = rng.randn(16000 * 2)
    for snr_db in (0.0, 10.0, 20.0):
        y = x + rng.randn(x.size) * 10 ** (-snr_db / 20)
        assert abs(M.sdr(x, y) - snr_db) < 0.5                           # white noise: nothing a 512-tap filter can undo
    assert abs(M.sdr(x, 3.0 * x) - 50.0) < 1e-6                          # clamp_db = 50
    delayed = np.concatenate([np.zeros(100), x[:-100]])
    assert M.sdr(x, delayed) > 20.0 > M.sdr(x, np.roll(x, 2000)) + 15.0    # a delay < 512 taps is an allowed distortion, a longer one is not
    # dense least-squares projection onto the 512 delayed copies on a short signal
    clean, noisy = _speechlike(6000, 16000, 5)
    X = np.stack([np.concatenate([np.zeros(k), clean[: clean.size - k]]) for k in range(512)], 1)
    p = X @ np.linalg.lstsq(X, noisy, rcond=None)[0]
    dense = 10 * np.log10((p ** 2).sum() / ((noisy - p) ** 2).sum())
    assert abs(M.sdr(clean, noisy) - dense) < 0.5
